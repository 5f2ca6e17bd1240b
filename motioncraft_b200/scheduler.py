"""Time schedules of the RePaint-style long-form sampler (mogen/models/utils/scheduler.py).

Only `get_schedule_jump_cjm_ddim` (scheduler.py:178-208) is reachable from the sampling path (the harmonising DDIM loop,
gaussian_diffusion.py:1079-1084); it is restated here for the product path -- the oracle keeps its own copy.
"""


def get_schedule_jump_cjm_ddim(time_respacing=25, jump_length=1, jump_n_sample=1):
    """Respaced timesteps the harmonising loop visits, ending with -1.

    The walk starts at `int(0.6 * time_respacing) - 1` (14 for the special case 25), goes down one step at a time and,
    on reaching a multiple of `jump_length` below `start - jump_length`, climbs back `jump_length` steps -- `jump_n_sample
    - 1` times per such step -- so every stretch is denoised `jump_n_sample` times with re-noising in between."""
    top = 15 if time_respacing == 25 else int(time_respacing * 0.6)
    remaining = {j: jump_n_sample - 1 for j in range(0, top - jump_length, jump_length)}
    seq, cur = [], top
    while cur >= 1:
        cur -= 1
        seq.append(cur)
        if remaining.get(cur, 0) > 0:
            remaining[cur] -= 1
            seq.extend(range(cur + 1, cur + jump_length + 1))
            cur += jump_length
    seq.append(-1)
    _check(seq, top)
    return seq


def _check(seq, top):
    """scheduler.py:47-61 (_check_times): steps of exactly one, strictly inside [-1, top]."""
    assert seq[0] > seq[1], (seq[0], seq[1])
    assert seq[-1] == -1, seq[-1]
    for a, b in zip(seq[:-1], seq[1:]):
        assert abs(a - b) == 1, (a, b)
    for t in seq:
        assert -1 <= t <= top, (t, top)


def count_draws(times, n_steps):
    """randn_like draws the reference makes along `times` (None: the plain n_steps loop): two per denoise call, one per undo."""
    if times is None:
        return 2 * n_steps
    n_den = sum(1 for a, b in zip(times[:-1], times[1:]) if b < a)
    return 2 * n_den + (len(times) - 1 - n_den)
