"""CPU models of three index / ordering schemes the decode kernels rely on (csrc/decode_stream.cu, csrc/sample.cu).  They restate the
device code's integer logic in numpy / Python so that the invariants hold independently of a GPU run; the `-m gpu` tests check the kernels."""
import numpy as np
import pytest


def test_mma_fragment_k_relabelling():
    """dec_linear_kernel feeds mma.sync.m16n8k16 from one 16-byte piece per row: thread (g, t) holds elements 8t..8t+7 of a 32-wide K block
    and hands words (2s, 2s+1) to k16 step s as the slot pairs (2t, 2t+1) / (2t+8, 2t+9) of BOTH operands.  Any bijection slot -> k that is
    the same for A and B leaves the product unchanged: the two steps must sum to x . W^T over the whole block."""
    rng = np.random.default_rng(0)
    X, W = rng.standard_normal((16, 32)), rng.standard_normal((8, 32))
    acc = np.zeros((16, 8))
    for s in range(2):
        A, B = np.zeros((16, 16)), np.zeros((16, 8))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            lo, hi, w = X[g, 8 * t:8 * t + 8], X[g + 8, 8 * t:8 * t + 8], W[g, 8 * t:8 * t + 8]
            word = lambda v, i: v[2 * i:2 * i + 2]
            # PTX fragment layout of m16n8k16 (row.col): a0 (g, 2t..), a1 (g+8, 2t..), a2 (g, 2t+8..), a3 (g+8, 2t+8..); b0 (2t.., g), b1 (2t+8.., g)
            A[g, 2 * t:2 * t + 2], A[g + 8, 2 * t:2 * t + 2] = word(lo, 2 * s), word(hi, 2 * s)
            A[g, 2 * t + 8:2 * t + 10], A[g + 8, 2 * t + 8:2 * t + 10] = word(lo, 2 * s + 1), word(hi, 2 * s + 1)
            B[2 * t:2 * t + 2, g], B[2 * t + 8:2 * t + 10, g] = word(w, 2 * s), word(w, 2 * s + 1)
        acc += A @ B
    assert np.abs(acc - X @ W.T).max() < 1e-12


def _before(va, ia, vb, ib):          # total order of sample.cu: value descending, index ascending
    return va > vb or (va == vb and ia < ib)


@pytest.mark.parametrize('NP', [32, 64, 256, 2048])
def test_bitonic_network_with_warp_local_stages(NP):
    """sample.cu's compare-exchange enumeration: exchange c of stride j pairs t = (c with a 0 inserted at bit log2 j) and t | j.  It sorts
    (ties broken by index), and for j <= 32 the 32 exchanges c in [32w, 32w+32) stay inside elements [64w, 64w+64): those stages need
    __syncwarp() only."""
    rng = np.random.default_rng(NP)
    val = list(rng.integers(0, 40, NP).astype(float))
    ref = sorted(range(NP), key=lambda i: (-val[i], i))
    idx = list(range(NP))
    k = 2
    while k <= NP:
        j = k >> 1
        while j > 0:
            for c in range(NP // 2):
                t = ((c & ~(j - 1)) << 1) | (c & (j - 1))
                p = t | j
                if j <= 32:
                    assert t // 64 == c // 32 and p // 64 == c // 32
                up = (t & k) == 0
                va, vb, ia, ib = val[t], val[p], idx[t], idx[p]
                if (_before(vb, ib, va, ia) if up else _before(va, ia, vb, ib)):
                    val[t], val[p], idx[t], idx[p] = vb, va, ib, ia
            j >>= 1
        k <<= 1
    assert idx == ref


@pytest.mark.parametrize('top_k', [1, 3, 8, 64])
def test_topk_selection_equals_sorted_prefix_with_ties(top_k):
    """The selection path of sample_block (repeated arg-max in the total order, continuing while the next maximum ties the k-th value) keeps
    exactly the entries HF's TopKLogitsWarper keeps (`scores >= k-th largest`), in the order of the full sort."""
    rng = np.random.default_rng(top_k)
    for trial in range(20):
        V = int(rng.integers(top_k, 200))
        val = list(rng.integers(0, 12, V).astype(float))       # many ties
        order = sorted(range(V), key=lambda i: (-val[i], i))
        kth = val[order[min(top_k, V) - 1]]
        want = [i for i in order if val[i] >= kth]
        taken, sel, kth_sel = set(), [], None
        r = 0
        while True:
            cand = [i for i in range(V) if i not in taken]
            if not cand:
                break
            wi = min(cand, key=lambda i: (-val[i], i))
            if r >= min(top_k, V) and val[wi] != kth_sel:
                break
            sel.append(wi)
            taken.add(wi)
            if r == min(top_k, V) - 1:
                kth_sel = val[wi]
            r += 1
        assert sel == want
