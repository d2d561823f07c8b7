"""Case table shared by make_golden.py (build container) and tests/test_oracle.py (everywhere)."""


def sd_checksum(sd):
    acc = 0.0
    for i, k in enumerate(sorted(sd)):
        acc += float(sd[k].double().sum()) * (1 + (i % 7)) + float(sd[k].double().abs().sum())
    return acc


CASES = [
    # name, in_channels, B, T, lengths, n_timesteps, solver, seed
    ("lj_full", 160, 2, 32, [32, 32], 3, "euler", 11),
    ("lj_ragged", 160, 3, 48, [48, 31, 17], 2, "euler", 12),
    ("vctk_ragged", 224, 2, 24, [24, 10], 2, "euler", 13),
    ("lj_midpoint", 160, 1, 16, [16], 2, "midpoint", 14),
    ("lj_long", 160, 1, 200, [200], 1, "euler", 15),
]
