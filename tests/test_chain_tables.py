"""CPU checks of two claims the second version of the brick gradient chain rests on (mf-lbm-cuda_b200/csrc/kernels_chain.cuh).

1. wsum_fill / wsum_index: the divisor of the weighted means of extrapolate_phi_toSolid and extrapolateNormalToSolid
   (/root/reference/src/main_iteration_GPU.cu:732-755, :880-906) - the weights of the contributing neighbours accumulated in the
   order q = 1 .. 18 - depends on the 18-bit neighbour mask only through (number of axis neighbours, number of diagonal
   neighbours), bit for bit, in both precisions: the kernels read it from a 91-entry table.
2. k_act_verdict / k_chain_extrap_cn_flat: list slot and entry offset of a brick are handed out by ONE atomic on a packed
   {bricks, entries} counter, so the offsets are non-decreasing in the slot whatever order the warps arrive in, and the brick of
   a flat entry number is the last slot whose offset is <= it (empty bricks are never selected).
"""
import numpy as np
import pytest


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_weight_sum_depends_on_neighbour_counts_only(dtype):
    w = [None] + [dtype(1) / dtype(18)] * 6 + [dtype(1) / dtype(36)] * 12   # includes/Module.h:114-122, computed in T
    masks = np.arange(1 << 18, dtype=np.int64)
    acc = np.zeros(masks.shape, dtype)
    for q in range(1, 19):   # the kernels' (and the reference's) accumulation order
        bit = (masks >> (q - 1)) & 1 == 1
        acc = np.where(bit, (acc + w[q]).astype(dtype), acc)
    table = np.zeros((7, 13), dtype)   # wsum_fill: n6 times 1/18, then n12 times 1/36
    for n6 in range(7):
        for n12 in range(13):
            s = dtype(0)
            for _ in range(n6):
                s = dtype(s + w[1])
            for _ in range(n12):
                s = dtype(s + w[7])
            table[n6, n12] = s
    popc = lambda x: np.array([bin(int(v)).count("1") for v in range(1 << 12)], np.int64)[x]
    n6 = popc(masks & 0x3f)
    n12 = popc(masks >> 6)
    assert np.array_equal(acc, table[n6, n12])
    assert acc.dtype == dtype


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_flat_entry_numbering_is_monotone_and_bisection_finds_the_brick(seed):
    rng = np.random.default_rng(seed)
    nbricks = 5000
    proc = rng.random(nbricks) < 0.3
    entries = np.where(rng.random(nbricks) < 0.5, 0, rng.integers(1, 70, nbricks))   # sb_start[b + 1] - sb_start[b]
    warps = [np.arange(w, min(w + 32, nbricks)) for w in range(0, nbricks, 32)]
    order = rng.permutation(len(warps))   # the order in which the warps' atomics arrive
    counter = 0                           # low 32 bits: bricks, high 32 bits: entries
    active, ent_off = {}, {}
    for wi in order:
        lanes = [b for b in warps[wi] if proc[b]]
        if not lanes:
            continue
        tot = int(sum(entries[b] for b in lanes))
        old = counter
        counter += (tot << 32) | len(lanes)   # ONE atomicAdd
        base, ebase = old & 0xffffffff, old >> 32
        run = 0
        for k, b in enumerate(lanes):   # lane order inside the warp
            active[base + k] = b
            ent_off[base + k] = ebase + run
            run += int(entries[b])
    n_active, total = counter & 0xffffffff, counter >> 32
    assert n_active == int(proc.sum()) and total == int(entries[proc].sum())
    off = np.array([ent_off[s] for s in range(n_active)])
    assert off[0] == 0 and np.all(np.diff(off) >= 0)
    seen = np.zeros(nbricks, np.int64)
    for g in range(total):   # k_chain_extrap_cn_flat
        lo, hi = 0, n_active - 1
        while lo < hi:
            mid = (lo + hi + 1) >> 1
            if off[mid] <= g:
                lo = mid
            else:
                hi = mid - 1
        b = active[lo]
        local = g - off[lo]
        assert 0 <= local < entries[b]
        seen[b] += 1
    assert np.array_equal(seen, np.where(proc, entries, 0))   # every entry of every listed brick exactly once


def test_python_emulation_mirrors_the_kernel_header():
    """the checks above must be about what the CUDA code does: same table shape, same index, same packed counter"""
    import re
    from pathlib import Path
    src = (Path(__file__).resolve().parent.parent / "mf-lbm-cuda_b200" / "csrc" / "kernels_chain.cuh").read_text()
    assert re.search(r"constexpr int WSUM_N = 7 \* 13;", src)
    assert re.search(r"return 13 \* __popc\(m & 0x3f\) \+ __popc\(m >> 6\);", src)
    assert re.search(r"for \(int n = 0; n < n6; n\+\+\) w \+= w_equ<T>\(1\);\s*for \(int n = 0; n < n12; n\+\+\) w \+= w_equ<T>\(7\);", src)
    assert re.search(r"\(\(unsigned long long\)\(unsigned\)tot << 32\) \| \(unsigned\)__popc\(vote\)", src)
    assert re.search(r"if \(\(unsigned\)__ldg\(ent_off \+ mid\) <= g\) lo = mid; else hi = mid - 1;", src)
