"""Seeded work lists for the RDO bit counter (xb200_rdo_bits): random coder states and syntax items that cover every
branch of src_base/xeve_mode.c:57-302 (skip / inter / mvp / per-component counters, P and B slices, forced-zero cbf
combinations, dense and sparse coefficient planes, the exp-golomb mvd extremes)."""
import numpy as np

from xeve_b200 import api


def rand_states(rng, n):
    st = np.zeros(n, api.SBAC)
    st["range"] = rng.integers(8192, 16384, n) & ~1
    st["m"] = (rng.integers(1, 257, (n, api.CM_COUNT)) << 1) | rng.integers(0, 2, (n, api.CM_COUNT))
    return st


def rand_items(rng, n, n_in):
    """items read states [0, n_in) and write states n_in + i (unique slots, as the ABI requires)."""
    it = np.zeros(n, api.BITS_ITEM)
    coefs, off = [], 0
    for i in range(n):
        r = it[i]
        r["kind"] = rng.integers(0, 4)
        r["slice_type"] = rng.integers(0, 2)
        l2 = int(rng.integers(3, 7))
        r["log2_cuw"] = r["log2_cuh"] = l2
        is_b = r["slice_type"] == 0
        pidx = int(rng.choice([0, 1, 2, 4])) if is_b else 0
        r["pidx"], r["ch"] = pidx, rng.integers(0, 3)
        r["ctx_skip"], r["ctx_pred_mode"] = rng.integers(0, 2), rng.integers(0, 3)
        nr = rng.integers(1, 5, 2)
        r["num_refp"] = nr
        r["refi"] = [rng.integers(0, nr[0]) if pidx in (0, 2, 4) else -1, rng.integers(0, nr[1]) if pidx in (1, 2, 4) else -1]
        r["mvp_idx"], r["all_preds"] = rng.integers(0, 4, 2), rng.integers(0, 2)
        r["mvd"] = rng.choice([0, 1, -1, 2, -3, 7, -40, 300, -2047, 2047, -2048, 32767, -32768], (2, 2))
        r["state_in"] = rng.integers(0, n_in)
        r["state_out"] = n_in + i if rng.random() < 0.5 else -1
        ny = 1 << (2 * l2)
        c = np.zeros(ny * 3 // 2, np.int16)
        m = rng.random(c.size) < rng.choice([0, 0.01, 0.1, 0.5])
        c[m] = rng.choice([1, -1, 2, -2, 3, 5, -9, 40, -300], int(m.sum()))
        if rng.random() < 0.2:
            c[ny - 1] = 1  # last scan position coded: no `last` flag follows it
        nn = [np.count_nonzero(c[:ny]), np.count_nonzero(c[ny:ny + ny // 4]), np.count_nonzero(c[ny + ny // 4:])]
        if rng.random() < 0.3:
            nn[int(rng.integers(0, 3))] = 0  # plane signalled as zero (the forced-zero tests of pinter_residue_rdo)
        r["nnz"], r["coef_off"] = nn, off
        coefs.append(c)
        off += c.size
    return it, np.concatenate(coefs)


def work(seed=5, n=3000, n_in=32):
    rng = np.random.default_rng(seed)
    st = np.concatenate([rand_states(rng, n_in), np.zeros(n, api.SBAC)])
    it, coef = rand_items(rng, n, n_in)
    return it, st, coef
