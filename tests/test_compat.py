"""The `mogen` import-path shell (compat/mogen) and the mmcv stand-in (compat/shims/mmcv): what the reference's
UNCHANGED tools do between `mmcv.Config.fromfile` and `model(return_loss=False, **data)` (tools/test.py:63-111,
mogen/apis/test.py:13-33), run in a fresh interpreter whose PYTHONPATH is the one INTEGRATION.md prescribes.

CPU part: config files of the reference itself when a checkout is present (/root/reference here; skipped on the GPU box),
otherwise an equivalent config written by the test.  GPU part: the same flow down to x_0 and against the oracle."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HAVE_REF = os.path.isdir(os.path.join(REF, "mogen", "apis"))

CONFIG_BASE = """
data_keys = ['motion', 'motion_mask', 'motion_length']
meta_keys = ['text', 'token']
data = dict(samples_per_gpu=128, workers_per_gpu=1, test=dict(type='TextMotionDataset', dataset_name='motionx', test_mode=False))
"""
CONFIG = """
_base_ = ['base/ds.py']
max_seq_len = 60
latent_dim = 512
time_embed_dim = 2048
model = dict(type='MotionDiffusion',
             model=dict(type='MCMTransformer', input_feats=322, max_seq_len=max_seq_len, latent_dim=latent_dim,
                        time_embed_dim=time_embed_dim, num_layers=8,
                        sa_block_cfg=dict(type='EfficientSelfAttention', latent_dim=max_seq_len, num_heads=4, dropout=0,
                                          time_embed_dim=time_embed_dim),
                        ca_block_cfg=dict(type='EfficientCrossAttention', latent_dim=latent_dim, text_latent_dim=256,
                                          num_heads=4, dropout=0, time_embed_dim=time_embed_dim),
                        ffn_cfg=dict(latent_dim=latent_dim, ffn_dim=1024, dropout=0, time_embed_dim=time_embed_dim),
                        text_encoder=dict(pretrained_model='clip', latent_dim=256, num_layers=4, num_heads=4, ff_size=2048,
                                          dropout=0, use_text_proj=True)),
             loss_recon=dict(type='MSELoss', loss_weight=1, reduction='none'),
             diffusion_train=dict(beta_scheduler='linear', diffusion_steps=1000, model_mean_type='epsilon',
                                  model_var_type='fixed_small'),
             diffusion_test=dict(beta_scheduler='linear', diffusion_steps=1000, model_mean_type='epsilon',
                                 model_var_type='fixed_small', respace='15,15,8,6,6'),
             inference_type='ddim')
data = dict(samples_per_gpu=256)
"""


def _run(code, cwd, extra_env=None, timeout=900):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "compat"), os.path.join(ROOT, "compat", "shims"), ROOT])
    env["PYTHONDONTWRITEBYTECODE"] = "1"        # never write into the working directory (it may be the reference checkout)
    env.update(extra_env or {})
    import tempfile
    with tempfile.TemporaryDirectory() as td:   # the script lives OUTSIDE cwd: sys.path[0] must not be the checkout
        script = os.path.join(td, "compat_probe.py")
        with open(script, "w") as f:
            f.write(textwrap.dedent(code))
        r = subprocess.run([sys.executable, script], cwd=str(cwd), env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    return r.stdout


def _write_cfg(tmp_path):
    (tmp_path / "base").mkdir(exist_ok=True)
    (tmp_path / "base" / "ds.py").write_text(CONFIG_BASE)
    (tmp_path / "cfg.py").write_text(CONFIG)
    return str(tmp_path / "cfg.py")


def test_mmcv_config_and_dictaction(tmp_path):
    cfgfile = _write_cfg(tmp_path)
    out = _run(f"""
        import argparse, mmcv
        from mmcv import DictAction
        cfg = mmcv.Config.fromfile({cfgfile!r})
        assert cfg.model.type == 'MotionDiffusion' and cfg.model.model.sa_block_cfg.latent_dim == 60
        assert cfg.data.samples_per_gpu == 256 and cfg.data.workers_per_gpu == 1          # child overrides, base merges
        assert cfg.data.test.type == 'TextMotionDataset' and cfg.get('fp16', None) is None
        cfg.data.test.test_mode = True
        assert cfg['data']['test']['test_mode'] is True
        p = argparse.ArgumentParser(); p.add_argument('--cfg-options', nargs='+', action=DictAction)
        a = p.parse_args(['--cfg-options', 'model.inference_type=ddpm', 'data.samples_per_gpu=8', 'x.y=[1,2]', 'z=true'])
        cfg.merge_from_dict(a.cfg_options)
        assert cfg.model.inference_type == 'ddpm' and cfg.data.samples_per_gpu == 8 and cfg.x.y == [1, 2] and cfg.z is True
        assert cfg.model.model.num_layers == 8                                           # untouched siblings survive
        cfg.model['opt'] = argparse.Namespace(overlap_len=0)
        assert cfg.model.opt.overlap_len == 0
        print('OK')
        """, tmp_path)
    assert "OK" in out


def test_registry_checkpoint_and_dataparallel_plumbing(tmp_path):
    """build_architecture(cfg.model) under the reference's import paths, a `model.`-prefixed checkpoint through
    load_checkpoint, and MMDataParallel + collate + DataContainer(cpu_only) delivering `motion_metas` as list[dict]."""
    cfgfile = _write_cfg(tmp_path)
    out = _run(f"""
        import argparse, torch, mmcv
        from mmcv.parallel import MMDataParallel, DataContainer as DC, collate
        from mmcv.runner import load_checkpoint, get_dist_info
        from mogen.models import build_architecture
        from mogen.models.builder import build_submodule, MODELS
        from mogen.models.transformers.controlnet import ControlT2MHalf
        from mogen.models.transformers.controlnet_mcm import ControlT2MHalf_MCM
        from mogen.models.utils.gaussian_diffusion import SpacedDiffusion
        import mogen
        assert mogen.digit_version('1.7.0') == (1, 7, 0, 0, 0, 0)
        cfg = mmcv.Config.fromfile({cfgfile!r})
        cfg.model['opt'] = argparse.Namespace(overlap_len=0)
        model = build_architecture(cfg.model)
        assert type(model).__name__ == 'MotionDiffusion' and type(model.model).__name__ == 'MCMTransformer'
        assert isinstance(model.diffusion_test, SpacedDiffusion) and model.diffusion_test.num_timesteps == 50
        assert build_submodule(None) is None and 'ControlT2MHalf_MCM' in MODELS
        # a checkpoint as mmcv's runner writes it: {{'meta': ..., 'state_dict': {{'model.<key>': tensor}}}}, saved from a DataParallel wrapper
        sd = {{'module.' + k: torch.full_like(v, 0.25) for k, v in model.state_dict().items()}}
        assert all(k.startswith('module.model.') for k in sd)
        torch.save({{'meta': {{}}, 'state_dict': sd}}, {str(tmp_path / 'ckpt.pth')!r})
        ck = load_checkpoint(model, {str(tmp_path / 'ckpt.pth')!r}, map_location='cpu')
        assert 'state_dict' in ck
        assert float(model.model.joint_embed.weight.mean()) == 0.25 and float(model.model.out.bias.mean()) == 0.25
        assert get_dist_info() == (0, 1)
        # the data path of tools/test.py: Collect wraps metas in DataContainer(cpu_only=True); collate; MMDataParallel scatters
        samples = [dict(motion=torch.zeros(60, 322), motion_mask=torch.ones(60), motion_length=torch.tensor(60),
                        motion_metas=DC(dict(text='t%d' % i, token=None), cpu_only=True)) for i in range(3)]
        batch = collate(samples, samples_per_gpu=3)
        assert batch['motion'].shape == (3, 60, 322) and isinstance(batch['motion_metas'], DC)
        seen = {{}}
        class Probe(torch.nn.Module):
            def forward(self, return_loss=False, **kw):
                seen.update(kw); return [0] * kw['motion'].shape[0]
        wrapped = MMDataParallel(Probe(), device_ids=[])            # CPU: scatter only unwraps
        assert len(wrapped(return_loss=False, **batch)) == 3
        assert seen['motion_metas'] == [dict(text='t0', token=None), dict(text='t1', token=None), dict(text='t2', token=None)]
        assert torch.equal(seen['motion_length'], torch.tensor([60, 60, 60]))
        # m2d_test.py:372-381: wrap the denoiser in the control net, re-assign architecture.model, load a checkpoint
        ccfg = mmcv.Config(dict(model=cfg.model, copy_blocks_num=2, control_cond_feats=35,
                                condition_encode_cfg=dict(dataset_name='finedance', condition_pre_encode=False, condition_cfg=True)))
        ctrl = ControlT2MHalf_MCM(model.model, copy_blocks_num=ccfg.copy_blocks_num, control_cond_feats=ccfg.control_cond_feats, cfg=ccfg).train()
        model.model = ctrl
        sd = {{k: torch.full_like(v, 0.5) for k, v in model.state_dict().items() if v.is_floating_point()}}
        torch.save({{'state_dict': sd}}, {str(tmp_path / 'ckpt2.pth')!r})
        load_checkpoint(model, {str(tmp_path / 'ckpt2.pth')!r}, map_location='cpu')
        assert float(model.model.controlnet[1].after_proj.weight.mean()) == 0.5
        assert float(model.model.base_model.out.weight.mean()) == 0.5
        try:
            ControlT2MHalf(None)
        except Exception as e:
            assert 'STMoGen' in str(e)
        print('OK')
        """, tmp_path)
    assert "OK" in out


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference checkout (build container only)")
def test_reference_configs_and_unchanged_api_loop(tmp_path):
    """With an unmodified reference checkout as the working directory (how its tools are launched): the reference's own
    configs/mcm/*.py build through `mogen.models.build_architecture`, `mogen.core` / `mogen.utils` / evaluator models come
    from the REFERENCE's files, and the reference's single_gpu_test loop (mogen/apis/test.py:13-33, loaded from its file)
    drives the wrapped B200 model up to the device boundary, where -- on this GPU-less host -- it fails loudly."""
    out = _run(r"""
        import argparse, importlib.util, sys, types, torch, mmcv
        from mmcv.parallel import MMDataParallel, DataContainer as DC, collate
        import mogen
        assert mogen.__path__[0].endswith('compat/mogen') and mogen.__path__[1] == '/root/reference/mogen'
        from mogen.models import build_architecture
        from mogen.models.transformers.controlnet_mcm import ControlT2MHalf_MCM
        import mogen.core.evaluation.utils as ev, mogen.utils as mu, mogen.models.rnns as rn, mogen.models.utils.word_vectorizer as q
        for m in (ev, mu, rn, q):
            assert m.__file__.startswith('/root/reference/'), m.__file__
        import mogen.models.transformers.mcm as mcm_mod
        assert 'motioncraft_b200' in mcm_mod.MCMTransformer.__module__
        for name, n_ctrl in (('mcm_t2m_smplx', 0), ('mcm_m2d_finedance', 4), ('mcm_s2g_beats2', 2)):
            cfg = mmcv.Config.fromfile('configs/mcm/%s.py' % name)
            cfg.data.test.test_mode = True
            cfg.model['opt'] = argparse.Namespace(overlap_len=0)
            model = build_architecture(cfg.model)
            assert type(model.model).__name__ == 'MCMTransformer' and len(model.model.temporal_decoder_blocks) == 8
            assert model.model.max_seq_len == cfg.model.model.max_seq_len
            if 'copy_blocks_num' in cfg:
                net = ControlT2MHalf_MCM(model.model, copy_blocks_num=cfg.copy_blocks_num,
                                         control_cond_feats=cfg.control_cond_feats, cfg=cfg).train()
                assert len(net.controlnet) == cfg.copy_blocks_num
                model.model = net
            print(name, 'built', sum(p.numel() for p in model.parameters()) // 10 ** 6, 'M params')
        # the reference's test loop, from its own file (its package __init__ also pulls in the training stack and
        # pytorch3d, which this image lacks)
        spec = importlib.util.spec_from_file_location('ref_apis_test', '/root/reference/mogen/apis/test.py')
        ref_test = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_test)
        cfg = mmcv.Config.fromfile('configs/mcm/mcm_t2m_smplx.py')
        cfg.model['opt'] = argparse.Namespace(overlap_len=0)
        model = MMDataParallel(build_architecture(cfg.model), device_ids=[])
        class DS(torch.utils.data.Dataset):
            def __len__(self): return 2
            def __getitem__(self, i):
                return dict(motion=torch.zeros(196, 322), motion_mask=torch.ones(196), motion_length=torch.tensor(196),
                            xf_proj=torch.zeros(2048), xf_out=torch.zeros(77, 256),
                            motion_metas=DC(dict(text='walk'), cpu_only=True))
        loader = torch.utils.data.DataLoader(DS(), batch_size=2, collate_fn=lambda b: collate(b, samples_per_gpu=2))
        from motioncraft_b200._lib import McmError
        try:
            ref_test.single_gpu_test(model, loader)
            raise SystemExit('expected the device boundary to refuse a CPU run')
        except McmError as e:
            assert 'CUDA' in str(e) or 'sm_100a' in str(e), str(e)
        print('OK')
        """, REF)
    assert "OK" in out


@pytest.mark.gpu
def test_unchanged_tool_flow_on_gpu_vs_oracle(tmp_path):
    """tools/test.py:63-111 + mogen/apis/test.py:13-33 against the B200 model on the device: Config.fromfile ->
    build_architecture -> load_checkpoint (`model.`-prefixed keys) -> MMDataParallel(device_ids=[0]) -> the collated
    batch with DataContainer metas -> list of per-sample dicts; x_0 against the CPU oracle."""
    from tests import common as C
    cfgfile = _write_cfg(tmp_path)
    T, B = 60, 2
    sd = C.base_state(T)
    torch.save({"meta": {}, "state_dict": {"model." + k: v for k, v in sd.items()}}, str(tmp_path / "ckpt.pth"))
    x, xf_out, xf_proj = C.inputs(B, T)
    torch.save(dict(x=x, xf_out=xf_out, xf_proj=xf_proj), str(tmp_path / "inputs.pt"))
    _run(f"""
        import argparse, torch, mmcv
        from mmcv.parallel import MMDataParallel, DataContainer as DC, collate
        from mmcv.runner import load_checkpoint
        from mogen.models import build_architecture
        cfg = mmcv.Config.fromfile({cfgfile!r})
        cfg.data.test.test_mode = True
        cfg.model['opt'] = argparse.Namespace(overlap_len=0)
        model = build_architecture(cfg.model)
        load_checkpoint(model, {str(tmp_path / 'ckpt.pth')!r}, map_location='cpu')
        model = MMDataParallel(model, device_ids=[0])
        model.eval()
        inp = torch.load({str(tmp_path / 'inputs.pt')!r})
        samples = [dict(motion=torch.zeros({T}, 322), motion_mask=torch.ones({T}), motion_length=torch.tensor({T}),
                        xf_proj=inp['xf_proj'][i], xf_out=inp['xf_out'][i],
                        motion_metas=DC(dict(text='clip %d' % i), cpu_only=True)) for i in range({B})]
        data = collate(samples, samples_per_gpu={B})
        data['inference_kwargs'] = dict(noise=inp['x'].cuda())
        with torch.no_grad():
            result = model(return_loss=False, **data)
        assert isinstance(result, list) and len(result) == {B} and result[1]['text'] == 'clip 1'
        assert result[0]['pred_motion'].device.type == 'cpu'
        torch.save(torch.stack([r['pred_motion'] for r in result]), {str(tmp_path / 'x0.pt')!r})
        """, tmp_path)
    got = torch.load(str(tmp_path / "x0.pt"))
    want = C.oracle_ddim(sd, x, xf_proj, xf_out)
    assert C.rel_l2(got, want) < 1e-3
