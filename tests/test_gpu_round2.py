"""GPU: parity at the configurations bench.py reports ("What's weak" 1-4) and the round-2 C-ABI entries.

Tolerances (fp32 reference, BASELINE.json north_star: 1e-3 relative):
    TOL_FAST   1e-3  relative L2 of the shipped fp16-operand path, POOLED and for the WORST SAMPLE
    TOL_MAX    3e-3  max-norm  max|a - b| / max|b|  (a single element may sit a few sigma out)
    TOL_PRECISE 5e-5 precise_all (every GEMM 3-pass bf16x2): the on-device bridge between the CPU oracle (which can only
                     afford a few samples of the full-size batch) and all 256 rows of it
"""
import os
import subprocess
import sys

import pytest
import torch

import motioncraft_b200 as M
from motioncraft_b200 import modules, synth
from motioncraft_b200._lib import McmError
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from oracle import mcm_oracle as O
from tests import common as C

pytestmark = pytest.mark.gpu
TOL_FAST, TOL_MAX, TOL_PRECISE = 1e-3, 3e-3, 5e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def per_sample_rel_l2(a, b):
    a, b = a.double().cpu().flatten(1), b.double().cpu().flatten(1)
    return ((a - b).norm(dim=1) / b.norm(dim=1))


def check_all_norms(got, want, tol=TOL_FAST, tol_max=TOL_MAX, what=""):
    pooled = C.rel_l2(got, want)
    worst = float(per_sample_rel_l2(got, want).max())
    mx = C.max_rel(got, want)
    assert pooled < tol and worst < tol and mx < tol_max, (what, pooled, worst, mx)
    return pooled, worst, mx


@pytest.mark.parametrize("name,T,n_ctrl,c_feats,c_len,B", [("s2g", 300, 2, 2048, 297, 2), ("m2d", 1024, 4, 35, 1024, 2)])
def test_ddim50_control_benchmark_configs_vs_oracle(name, T, n_ctrl, c_feats, c_len, B):
    """BASELINE configs 2 / 3 (speech-to-gesture T=300 + 2 control blocks, music-to-dance T=1024 + 4 control blocks): the
    whole 50-step DDIM trajectory -- the compounding the single-forward test cannot see; at T=1024 the channel attention is
    63 % of the FLOPs with K=1024 fp16 GEMMs -- against the fp32 CPU oracle (controlnet_mcm.py:306-361 under
    gaussian_diffusion.py:925-1049)."""
    sd = synth.synth_state_dict(modules.ctrl_state_shapes(T, n_ctrl, c_feats))
    x, xf_out, xf_proj = C.inputs(B, T)
    c = synth.synth_tensor("c", (B, c_len, c_feats), synth.SEED_C_EMB)
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    with torch.no_grad():
        want = O.ddim_sample_loop(lambda xx, tt: O.control_forward(sd, xx, tt, xf_proj, xf_out, c), x, tables, tmap)
    eng = DenoiserEngine(modules.engine_state_from_ctrl(sd), seq_len=T, max_batch=B, num_ctrl_blocks=n_ctrl,
                         ctrl_cond_feats=c_feats)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda(), c.cuda())
    x0 = eng.sample(SamplerTables(tables, tmap, "ddim"), x.cuda())
    print(name, "pooled / worst-sample / max-norm:", check_all_norms(x0, want, what=name))
    eng.close()


def test_ddpm_200_steps_T196_vs_oracle():
    """Long ancestral chain: 200 DDPM steps (SpacedDiffusion '200') with scripted per-step noise at T=196; the shipped
    mcm_t2m config samples with DDPM (configs/mcm/mcm_t2m_smplx.py:73-79).  gaussian_diffusion.py:634-797."""
    T, B, n = 196, 1, 200
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T)
    noise = synth.synth_tensor("step_noise200", (n, B, T, 322), synth.SEED_STEP_NOISE)
    want = C.oracle_ddpm(sd, x, xf_proj, xf_out, noise, respace=str(n))
    tables, tmap = O.spaced_tables(1000, str(n))
    assert len(tmap) == n
    eng = DenoiserEngine(C.hot(sd), seq_len=T, max_batch=B)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    st = SamplerTables(tables, tmap, "ddpm")
    x0 = eng.sample(st, x.cuda(), noise.cuda())
    print("ddpm-200 pooled / worst / max:", check_all_norms(x0, want, what="ddpm200"))
    # the host-buffer entry with HOST noise (copied one step at a time) returns the same bits
    x0h = eng.sample_host(st, x.clone().pin_memory(), None, noise.clone().pin_memory())
    assert torch.equal(x0h, x0.cpu())
    eng.close()


def test_full_batch_worst_sample_and_max_norm_t2m():
    """BASELINE config 1 at FULL size (B=256, T=196, 50-step DDIM), every row checked:
      * 8 rows spread over the batch against the fp32 CPU oracle -- pooled, worst-sample and max-norm;
      * ALL 256 rows of the shipped fp16 path against the precise_all engine (3-pass bf16x2 GEMMs), which itself is held
        to 5e-5 of the oracle on those 8 rows: worst-sample rel-L2 and max-norm over the whole batch."""
    T, B = 196, 256
    sd = C.base_state(T)
    x = synth.synth_rows("x_T", (T, 322), synth.SEED_XT, 0, B)
    xf_out = synth.synth_rows("xf_out", (77, 256), synth.SEED_XF_OUT, 0, B)
    xf_proj = synth.synth_rows("xf_proj", (2048,), synth.SEED_XF_PROJ, 0, B)
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    st = SamplerTables(tables, tmap, "ddim")
    sel = [0, 37, 64, 101, 128, 170, 222, 255]
    want = C.oracle_ddim(sd, x[sel], xf_proj[sel], xf_out[sel])

    eng = DenoiserEngine(C.hot(sd), seq_len=T, max_batch=B)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    fast = eng.sample(st, x.cuda()).cpu()
    eng.close()
    print("fast vs oracle (8 rows):", check_all_norms(fast[sel], want, what="fast/oracle"))

    engp = DenoiserEngine(C.hot(sd), seq_len=T, max_batch=B, precise_all=True)
    engp.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    prec = engp.sample(st, x.cuda()).cpu()
    engp.close()
    print("precise vs oracle (8 rows):", check_all_norms(prec[sel], want, TOL_PRECISE, 3 * TOL_PRECISE, "precise/oracle"))
    ps = per_sample_rel_l2(fast, prec)
    print("fast vs precise, all 256 rows: worst sample %.3e (row %d), median %.3e, max-norm %.3e"
          % (float(ps.max()), int(ps.argmax()), float(ps.median()), C.max_rel(fast, prec)))
    check_all_norms(fast, prec, what="fast/precise all rows")


def test_device_generated_noise():
    """Stochastic samplers without explicit noise: the library draws each step's noise on the device (Philox4x32-10 +
    Box-Muller keyed by (seed, step)) into one step-sized buffer.  Reproducible from the seed, different across seeds and
    steps, and standard normal."""
    T, B = 60, 4
    x, xf_out, xf_proj = C.inputs(B, T)
    eng = DenoiserEngine(C.hot(C.base_state(T, 2)), seq_len=T, max_batch=B, num_layers=2)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    tables, tmap = O.spaced_tables(1000, "10")
    a = eng.sample(SamplerTables(tables, tmap, "ddpm", seed=7), x.cuda())
    b = eng.sample(SamplerTables(tables, tmap, "ddpm", seed=7), x.cuda())
    c = eng.sample(SamplerTables(tables, tmap, "ddpm", seed=8), x.cuda())
    assert torch.equal(a, b) and not torch.equal(a, c) and torch.isfinite(a).all()
    # through the reference-facing front end: torch.manual_seed controls the run
    from motioncraft_b200 import diffusion
    d = diffusion.build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon",
                                       model_var_type="fixed_small", respace="10"))
    net = M.MCMTransformer(**modules.mcm_config(T, num_layers=2))
    net.use_text_proj = True
    net.load_state_dict(C.base_state(T, 2))
    net = net.cuda().eval()
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    torch.manual_seed(5)
    r1 = d.p_sample_loop(net, (B, T, 322), clip_denoised=False, model_kwargs=kw)
    torch.manual_seed(5)
    r2 = d.p_sample_loop(net, (B, T, 322), clip_denoised=False, model_kwargs=kw)
    assert torch.equal(r1, r2)
    # the generator itself: moments of 4 M draws, independence of consecutive streams
    import ctypes
    n = 1 << 22
    z = torch.empty(2, n, device="cuda")
    for sub in (0, 1):
        assert eng.lib.mcm_test_randn(ctypes.c_void_p(z[sub].data_ptr()), n, 1234, sub, None) == 0
    torch.cuda.synchronize()
    zz = z.double()
    assert abs(float(zz.mean())) < 3e-3 and abs(float(zz.std()) - 1.0) < 3e-3
    assert abs(float((zz ** 3).mean())) < 1e-2 and abs(float((zz ** 4).mean()) - 3.0) < 3e-2     # skewness, kurtosis
    assert abs(float((zz[0] * zz[1]).mean())) < 3e-3                                             # streams uncorrelated
    assert abs(float((zz[0, :-1] * zz[0, 1:]).mean())) < 3e-3                                    # lag-1 autocorrelation
    assert float(zz.abs().max()) > 4.5 and torch.isfinite(z).all()                               # tails are populated
    eng.close()


def test_refinalize_in_place_and_forward_test_entry():
    """mcm_finalize_params may be called again (weights re-packed in place, include/mcm_b200.h:14) and
    MCMTransformer.forward_test(h, src_mask, emb, xf_out) (mcm.py:93-102) runs through mcm_layers_forward."""
    T, B = 60, 3
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T)
    eng = DenoiserEngine(C.hot(sd), seq_len=T, max_batch=B)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    e1 = eng.denoise(x.cuda(), 300)
    sd2 = {k: v * 0.5 for k, v in sd.items()}
    eng.load_params(C.hot(sd2))                                  # second finalize on the same context
    with pytest.raises(McmError):
        eng.denoise(x.cuda(), 300)                               # conditions depend on the weights: must be re-prepared
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    e2 = eng.denoise(x.cuda(), 300)
    assert C.rel_l2(e2, C.oracle_forward(sd2, x, 300, xf_proj, xf_out)) < TOL_FAST
    eng.load_params(C.hot(sd))
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    assert torch.equal(eng.denoise(x.cuda(), 300), e1)           # back to the first weights: the same bits
    eng.close()

    net = M.MCMTransformer(**modules.mcm_config(T))
    net.use_text_proj = True
    net.load_state_dict(sd)
    net = net.cuda().eval()
    col = {}
    want = C.oracle_forward(sd, x, 300, xf_proj, xf_out, collect=col)
    got = net.forward_test(h=col["h0"].cuda(), src_mask=torch.ones(B, T, 1).cuda(), emb=col["emb"].cuda(), xf_out=xf_out.cuda())
    assert C.rel_l2(got, want) < TOL_FAST
    eng0 = net._engine
    net.load_state_dict(sd2)                                     # the module re-packs in place: same engine object
    got2 = net.forward_test(h=col["h0"].cuda(), emb=col["emb"].cuda(), xf_out=xf_out.cuda())
    assert net._engine is eng0
    col2 = {}
    want2 = C.oracle_forward(sd2, x, 300, xf_proj, xf_out, collect=col2)
    # h0 / emb of the FIRST weights fed to the layers of the second: compare with the oracle evaluated the same way
    sd2_64 = {k: v.double() for k, v in sd2.items()}
    with torch.no_grad():
        hh = col["h0"].double()
        for i in range(8):
            hh = O.decoder_layer(hh, xf_out.double(), col["emb"].double(), sd2_64, f"temporal_decoder_blocks.{i}", 4)
        want2 = torch.nn.functional.linear(hh, sd2_64["out.weight"], sd2_64["out.bias"])
    assert C.rel_l2(got2, want2) < TOL_FAST


def test_condition_cache_never_serves_a_recycled_address():
    """the identity cache of prepare_conditions_cached must not hit when a NEW batch lands on the address of a
    freed one.  The cache entry now owns references to the tensors it was keyed on."""
    T, B = 60, 2
    eng = DenoiserEngine(C.hot(C.base_state(T, 2)), seq_len=T, max_batch=B, num_layers=2)
    x, xf_out, xf_proj = C.inputs(B, T)
    a, p = xf_out.cuda(), xf_proj.cuda()
    eng.prepare_conditions_cached(a, p)
    e_a = eng.denoise(x.cuda(), 100)
    ptr = a.data_ptr()
    del a                                        # without the strong reference the allocator would reuse this block ...
    b = (xf_out * -1.0).cuda()                   # ... for the next batch of the same shape
    assert b.data_ptr() != ptr                   # the cache keeps `a` alive, so the address cannot be recycled
    eng.prepare_conditions_cached(b, p)
    e_b = eng.denoise(x.cuda(), 100)
    assert not torch.equal(e_a, e_b)
    eng.close()


def test_scheduling_options_after_first_capture():
    """options set AFTER a step graph was captured must take effect (the graph key covers dual / chunk, and
    mcm_set_option drops stale graphs); results stay bit-identical."""
    T, B = 60, 5
    x, xf_out, xf_proj = C.inputs(B, T)
    eng = DenoiserEngine(C.hot(C.base_state(T, 2)), seq_len=T, max_batch=B, num_layers=2)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    tables, tmap = O.spaced_tables(1000, "10")
    st = SamplerTables(tables, tmap, "ddim")
    ref = eng.sample(st, x.cuda())
    from motioncraft_b200 import _lib
    n_dual = _lib.kernel_launches()
    eng.sample(st, x.cuda())
    n_dual = _lib.kernel_launches() - n_dual
    for name, val in (("dual", 0), ("chunk", 2), ("chunk", 0), ("dual", 1)):
        eng.set_option(name, val)
        n0 = _lib.kernel_launches()
        assert torch.equal(eng.sample(st, x.cuda()), ref), (name, val)
        n = _lib.kernel_launches() - n0
        if (name, val) == ("chunk", 2):
            assert n > n_dual                    # 3 passes through the layer stack instead of 2 halves: more launches
    eng.close()


_TWO_RANK = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from motioncraft_b200 import dist as mdist, modules, synth
from motioncraft_b200.engine import DenoiserEngine, SamplerTables
from oracle import mcm_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
T, N = 196, 24
sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T)).items() if ".ffn_channel." not in k}
tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
st = SamplerTables(tables, tmap, "ddim")
def run(lo, hi):
    eng = DenoiserEngine(sd, seq_len=T, max_batch=hi - lo, device=dev)
    eng.set_option("fused_min_rows", 0)     # one schedule for every launch size: results comparable bit for bit
    eng.prepare_conditions(synth.synth_rows("xf_out", (77, 256), synth.SEED_XF_OUT, lo, hi).to(dev),
                           synth.synth_rows("xf_proj", (2048,), synth.SEED_XF_PROJ, lo, hi).to(dev))
    out = eng.sample(st, synth.synth_rows("x_T", (T, 322), synth.SEED_XT, lo, hi).to(dev))
    eng.close()
    return out
lo, hi = mdist.shard_range(N, rank, world)
full = mdist.gather_rows(run(lo, hi), N)             # the ONE collective of the multi-GPU path (NCCL all-gather)
if rank == 0:
    single = run(0, N)                               # the same global rows on one GPU
    assert torch.equal(full, single), float((full - single).abs().max())
    print("2-rank gathered x0 == 1-GPU x0 (bit-equal), rows", N)
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_nccl_gather_is_bit_equal_to_single_gpu(tmp_path):
    """Multi-GPU result on hardware: 2 ranks sample their shards of a globally seeded batch, ONE NCCL all-gather
    reassembles x_0 (motioncraft_b200/dist.py, replacing mogen/apis/test.py:131-163), and the gathered tensor is
    torch.equal to the same batch sampled on one GPU."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK % ROOT)
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[:3000] + r.stderr[-1500:]
    assert "bit-equal" in r.stdout


def test_text_conditioning_stack_on_device(golden_dir):
    """text_pre_proj -> 4-layer encoder -> text_ln -> text_proj runs ONCE per sampling run on the
    device (torch library kernels) from `clip_feat` + the EOT position; against the reference's encode_text output, then the
    whole text -> x_0 chain through MotionDiffusion.forward against the oracle fed with the reference's embeddings."""
    import numpy as np
    g = np.load(os.path.join(golden_dir, "text_stack.npz"))
    T, B = 60, int(g["xf_proj"].shape[0])
    dt = dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon", model_var_type="fixed_small")
    cfg = dict(type="MotionDiffusion",
               model=dict(type="MCMTransformer", **modules.mcm_config(T, text_encoder=dict(modules.TEXT_ENCODER_CFG))),
               loss_recon=dict(type="MSELoss", loss_weight=1, reduction="none"), diffusion_train=dt,
               diffusion_test=dict(dt, respace="15,15,8,6,6"), inference_type="ddim")
    arch = M.build_architecture(cfg)
    sd = dict(C.base_state(T))
    sd.update(synth.synth_state_dict(modules.text_state_shapes()))
    arch.model.load_state_dict(sd)
    arch = arch.cuda().eval()
    clip_feat = synth.synth_tensor("clip_feat", (B, 77, 512), synth.SEED_CLIP_FEAT)
    eos = torch.from_numpy(g["eos_index"])
    cond = arch.model.get_precompute_condition(text=list(g["texts"]), clip_feat=clip_feat.cuda(), eos_index=eos, device="cuda")
    assert C.rel_l2(cond["xf_proj"], g["xf_proj"]) < 1e-5 and C.rel_l2(cond["xf_out"], g["xf_out"]) < 1e-5
    x = synth.synth_tensor("x_T", (B, T, 322), synth.SEED_XT)
    out = arch(motion=torch.zeros(B, T, 322).cuda(), motion_mask=torch.ones(B, T).cuda(),
               motion_length=torch.full((B,), T).cuda(), motion_metas=[{"text": t} for t in g["texts"]],
               clip_feat=clip_feat.cuda(), eos_index=eos, inference_kwargs={"noise": x.cuda()}, return_loss=False)
    got = torch.stack([o["pred_motion"] for o in out])
    want = C.oracle_ddim(C.base_state(T), x, torch.from_numpy(g["xf_proj"]), torch.from_numpy(g["xf_out"]))
    assert C.rel_l2(got, want) < TOL_FAST
    assert out[0]["text"] == str(g["texts"][0])


def test_longform_window_pipeline_vs_oracle():
    """SURVEY.md 8 row f-2, the rest: the sliding-window drivers of tools/m2d_test.py:145-222 / tools/s2g_test.py:144-241 as a
    device-side pipeline -- 3 overlapping windows of 2 sequences at once, window i + 1 pinned to the de-normalised tail of
    window i on the device, harmonising RePaint loop in every pinned window, scripted noise -- against the oracle's
    restatement of the tools' loop.  The pipeline must not wait for the GPU between windows."""
    import argparse
    import numpy as np
    from motioncraft_b200 import diffusion, longform
    T, B, W, pre, L = 60, 2, 3, 12, 10
    sd = synth.synth_state_dict(C.ctrl_shapes(T, 2, 35))
    base = M.MCMTransformer(**modules.mcm_config(T))
    base.use_text_proj = True
    cfg = dict(model=dict(model=modules.mcm_config(T)),
               condition_encode_cfg=dict(dataset_name="finedance", condition_pre_encode=False, condition_cfg=True))
    net = M.ControlT2MHalf_MCM(base, copy_blocks_num=2, control_cond_feats=35, cfg=cfg)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    opt = argparse.Namespace(no_repaint=False, same_overlap_noisy=False, addBlend=True, overlap_len=L, no_resample=False,
                             timestep_respacing="ddim50", jump_length=3, jump_n_sample=5)
    d = diffusion.build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon",
                                       model_var_type="fixed_small", respace="15,15,8,6,6"), opt=opt)
    rng = np.random.default_rng(3)
    mean, std = rng.standard_normal(322), 0.5 + rng.random(322)          # float64, as np.load of the datasets' statistics
    round_l = T - pre
    total = (W - 1) * round_l + T
    music = synth.synth_tensor("music", (B, total, 35), synth.SEED_C_M2D)
    _, xf_out, xf_proj = C.inputs(B, T)
    x_T = [synth.synth_tensor(f"x_T_w{i}", (B, T, 322), synth.SEED_XT) for i in range(W)]
    times = O.schedule_jump_cjm_ddim(50, 3, 5)
    n_draws = 2 * sum(1 for a, b in zip(times[:-1], times[1:]) if b < a) + sum(1 for a, b in zip(times[:-1], times[1:]) if b >= a)
    rnoise = [synth.synth_tensor(f"rp_w{i}", (n_draws, B, T, 322), synth.SEED_REPAINT_NOISE) for i in range(W)]
    win_c = [music[:, i * round_l: i * round_l + T] for i in range(W)]
    kwargs = [dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda(), c=win_c[i].cuda()) for i in range(W)]
    torch.cuda.synchronize()
    got = longform.sample_windows(net, d, kwargs, motion_length=T, pre_frames=pre, mean=mean, std=std,
                                  noise=[x.cuda() for x in x_T], repaint_noise=[r.cuda() for r in rnoise])
    done = torch.cuda.Event()
    done.record()
    still_running = not done.query()          # the host is back while the device still works through the windows
    torch.cuda.synchronize()
    assert got.shape == (B, total, 322) and got.dtype == torch.float64
    tables, tmap = O.spaced_tables(1000, "15,15,8,6,6")
    with torch.no_grad():
        want = O.longform_windows(lambda i: (lambda xx, tt: O.control_forward(sd, xx, tt, xf_proj, xf_out, win_c[i])), W, B, T,
                                  pre, L, tables, tmap, tables["betas"], mean, std, x_T,
                                  [[r[j] for j in range(r.shape[0])] for r in rnoise], times=times)
    assert C.rel_l2(got, want) < TOL_FAST
    assert still_running, "sample_windows blocked on the device"


def test_ddpm_1000_steps_shipped_default_vs_oracle():
    """configs/mcm/mcm_t2m_smplx.py:73-79 ships `inference_type='ddpm'` with the un-respaced 1000-step schedule: the full
    ancestral chain (T = 60, B = 1, scripted per-step noise) against the fp32 oracle, and
    the same run with noise generated on the device (finite, reproducible through torch.manual_seed)."""
    from motioncraft_b200 import diffusion
    T, B, n = 60, 1, 1000
    sd = C.base_state(T)
    x, xf_out, xf_proj = C.inputs(B, T)
    noise = synth.synth_tensor("step_noise1000", (n, B, T, 322), synth.SEED_STEP_NOISE)
    want = C.oracle_ddpm(sd, x, xf_proj, xf_out, noise, respace=None)
    net = M.MCMTransformer(**modules.mcm_config(T))
    net.use_text_proj = True
    net.load_state_dict(sd)
    net = net.cuda().eval()
    d = diffusion.build_diffusion(dict(beta_scheduler="linear", diffusion_steps=1000, model_mean_type="epsilon",
                                       model_var_type="fixed_small"))
    assert d.num_timesteps == n
    kw = dict(xf_proj=xf_proj.cuda(), xf_out=xf_out.cuda())
    x0 = d.p_sample_loop(net, (B, T, 322), noise=x.cuda(), clip_denoised=False, model_kwargs=kw, step_noise=noise.cuda())
    print("ddpm-1000 pooled / worst / max:", check_all_norms(x0, want, what="ddpm1000"))
    torch.manual_seed(11)
    a = d.p_sample_loop(net, (B, T, 322), noise=x.cuda(), clip_denoised=False, model_kwargs=kw)
    torch.manual_seed(11)
    b = d.p_sample_loop(net, (B, T, 322), noise=x.cuda(), clip_denoised=False, model_kwargs=kw)
    assert torch.isfinite(a).all() and torch.equal(a, b) and not torch.equal(a, x0)


_ROW_KERNEL_SCRIPT = r"""
import os, sys, torch
sys.path.insert(0, sys.argv[1])
from motioncraft_b200 import modules, synth
from motioncraft_b200.engine import DenoiserEngine
outs = {}
for T, B in ((196, 3), (300, 3), (1024, 2)):
    sd = {k: v for k, v in synth.synth_state_dict(modules.state_shapes(seq_len=T, num_layers=1)).items() if ".ffn_channel." not in k}
    eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_layers=1)
    eng.set_option("fused", 0); eng.set_option("fused_sa", 0); eng.set_option("dual", 0); eng.set_option("graph", 0)
    g = torch.Generator().manual_seed(T)
    h = torch.randn(B, T, 512, generator=g).cuda(); emb = torch.randn(B, 2048, generator=g).cuda()
    eng.prepare_conditions(torch.randn(B, 77, 256, generator=g).cuda(), torch.randn(B, 2048, generator=g).cuda())
    outs[str(T)] = eng.block_forward(0, 0, h, emb).cpu()
    eng.close()
torch.save(outs, sys.argv[2])
"""


def _run_row_kernel_script(tmp_path, tag, **env):
    script = tmp_path / "rows.py"
    script.write_text(_ROW_KERNEL_SCRIPT)
    out = tmp_path / f"rows_{tag}.pt"
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, str(script), ROOT, str(out)], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return torch.load(out)


def test_per_head_channel_attention_contractions_match_the_masked_product(tmp_path):
    """The shape-dependent kernel choice must not change results: the per-head channel-attention contractions (wide heads,
    T = 1024) match the masked full product (same products, the fp32 accumulation only skips exact zeros); narrow heads
    (T = 196, 300) do not take that path at all."""
    base = _run_row_kernel_script(tmp_path, "base")
    full = _run_row_kernel_script(tmp_path, "full", MCM_SA_PERHEAD=0)
    for k in base:
        assert torch.isfinite(base[k]).all()
    assert torch.equal(full["196"], base["196"]) and torch.equal(full["300"], base["300"])
    assert C.rel_l2(base["1024"], full["1024"]) < 1e-6


def test_hoisted_modulation_is_scheduling_only_and_survives_a_longer_schedule():
    """The sampler computes the timestep-conditioned AdaLN modulation of ALL steps before the loop (hoist_mod, default on) and
    every captured step gathers its slice from that table.  (1) x_0 is bit-identical with the hoist off; (2) a LONGER schedule
    on the same engine re-allocates the table -- captured step graphs must not keep gathering from the old one: the result
    equals a fresh engine's; (3) going back to the short schedule still reproduces the first result."""
    T, B = 60, 3
    x, xf_out, xf_proj = C.inputs(B, T)
    sd = C.hot(C.base_state(T, 2))

    def tables(respace):
        tb, tmap = O.spaced_tables(1000, respace)
        return SamplerTables(tb, tmap, "ddim")

    eng = DenoiserEngine(sd, seq_len=T, max_batch=B, num_layers=2)
    eng.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    short = eng.sample(tables("10"), x.cuda())
    long_ = eng.sample(tables("25"), x.cuda())                  # 25 x B rows > 10 x B: the table grows
    assert torch.equal(eng.sample(tables("10"), x.cuda()), short)
    eng.set_option("hoist_mod", 0)
    assert torch.equal(eng.sample(tables("10"), x.cuda()), short)
    assert torch.equal(eng.sample(tables("25"), x.cuda()), long_)
    eng.close()
    fresh = DenoiserEngine(sd, seq_len=T, max_batch=B, num_layers=2)
    fresh.prepare_conditions(xf_out.cuda(), xf_proj.cuda())
    assert torch.equal(fresh.sample(tables("25"), x.cuda()), long_)
    fresh.close()
