"""CPU model of the shared-memory layout / fragment index math of the two DMMA kernels (csrc/gram.cuh, csrc/syrk.cuh).

It replays, in numpy, exactly what one CTA does: TMA 128B-swizzled tile placement, the per-lane 128-bit fragment loads,
the m8n8k4 fragment ownership (PTX ISA: A[m=g][k=t], B[k=t][n=g], C[m=g][n=2t+{0,1}]), the accumulator -> (i, j) maps of
the epilogues, and checks (a) the result equals the plain matrix product and (b) every quarter-warp of every 128-bit
load touches 8 distinct 16-byte chunks modulo 128 bytes (= all 32 banks once: conflict-free).
No GPU needed; this pins the index math the CUDA code relies on.
"""
import numpy as np


def swz128(byte_off: int) -> int:
    """CU_TENSOR_MAP_SWIZZLE_128B: 16-byte chunk index (bits 4-6) ^= 128-byte row index mod 8 (bits 7-9)."""
    return byte_off ^ (((byte_off >> 7) & 7) << 4)


def tma_tile(src_rows: np.ndarray) -> np.ndarray:
    """rows x 16 doubles (128 B per row) -> flat smem image (doubles) with the 128B swizzle applied."""
    rows, cols = src_rows.shape
    assert cols == 16
    smem = np.full(rows * 16, np.nan)
    for r in range(rows):
        for c in range(16):
            smem[swz128(r * 128 + c * 8) // 8] = src_rows[r, c]
    return smem


def lds128(smem: np.ndarray, byte_off: int):
    assert byte_off % 16 == 0
    return smem[byte_off // 8], smem[byte_off // 8 + 1]


def check_conflict_free(addrs_by_lane):
    for q in range(4):
        chunks = {(addrs_by_lane[l] // 16) % 8 for l in range(8 * q, 8 * q + 8)}
        assert len(chunks) == 8, f"quarter-warp {q} is bank-conflicted: {sorted(chunks)}"


def dmma(acc, afrag, bfrag):
    """acc[lane] = [c0, c1]; afrag[lane], bfrag[lane] scalars. D = A(8x4) B(4x8) + C with the PTX fragment layout."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t] = afrag[lane]
        B[t, g] = bfrag[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane][0] += D[g, 2 * t]
        acc[lane][1] += D[g, 2 * t + 1]


def test_gram_fragment_mapping():
    rng = np.random.default_rng(1)
    BM = BN = 128
    BK = 16
    d = 16
    XA = rng.standard_normal((BM, d))
    XB = rng.standard_normal((BN, d))
    sA, sB = tma_tile(XA), tma_tile(XB)
    out = np.full((BM, BN), np.nan)
    for warp in range(8):
        wm, wn = warp >> 1, warp & 1
        acc = [[[[0.0, 0.0] for _ in range(32)] for _ in range(8)] for _ in range(4)]  # [mt][nt][lane][c]
        for h in range(2):
            a = [[None] * 32 for _ in range(4)]
            b = [[None] * 32 for _ in range(8)]
            for mt in range(4):
                addrs = []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    pg = ((g & 1) << 2) | (g >> 1)
                    xo = (t ^ pg) << 4
                    off = ((32 * wm + pg) * 128 + xo + mt * 1024) ^ (h << 6)
                    addrs.append(off)
                    a[mt][lane] = lds128(sA, off)
                check_conflict_free(addrs)
            for nt in range(8):
                addrs = []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    pg = ((g & 1) << 2) | (g >> 1)
                    xo = (t ^ pg) << 4
                    off = ((64 * wn + pg) * 128 + xo + nt * 1024) ^ (h << 6)
                    addrs.append(off)
                    b[nt][lane] = lds128(sB, off)
                check_conflict_free(addrs)
            for comp in range(2):
                for mt in range(4):
                    for nt in range(8):
                        dmma(acc[mt][nt], [a[mt][l][comp] for l in range(32)], [b[nt][l][comp] for l in range(32)])
        for mt in range(4):
            for nt in range(8):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    pg = ((g & 1) << 2) | (g >> 1)
                    for c in range(2):
                        i = 32 * wm + 8 * mt + pg
                        j = 64 * wn + 8 * nt + t + 4 * c
                        assert np.isnan(out[i, j])
                        out[i, j] = acc[mt][nt][lane][c]
    ref = XA @ XB.T
    assert np.allclose(out, ref, rtol=1e-13, atol=1e-13)


def test_syrk_fragment_mapping():
    rng = np.random.default_rng(2)
    BT, BK = 128, 16
    A = rng.standard_normal((BK, 2 * BT))  # 16 stage rows; columns 0..127 = i block, 128..255 = j block
    w = rng.uniform(0.5, 2.0, size=BK)
    # stage image: per operand 8 boxes of (16 rows x 16 cols), each 2 KB, 128B-swizzled
    def stage(cols0):
        return [tma_tile(A[:, cols0 + 16 * bx: cols0 + 16 * bx + 16]) for bx in range(8)]

    sI, sJ = stage(0), stage(BT)
    acc_all = {}
    for warp in range(8):
        wm, wn = warp >> 1, warp & 1
        acc = [[[[0.0, 0.0] for _ in range(32)] for _ in range(8)] for _ in range(4)]
        for h in range(2):
            for par in range(2):
                a = [[None] * 32 for _ in range(2)]
                b = [[None] * 32 for _ in range(4)]
                addrs = []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    rho = 8 * h + 2 * t + par
                    off = rho * 128 + ((g ^ (2 * t + par)) << 4)
                    addrs.append(off)
                    for bx in range(2):
                        x, y = lds128(sI[2 * wm + bx], off)
                        a[bx][lane] = (x * w[rho], y * w[rho])
                    for bx in range(4):
                        b[bx][lane] = lds128(sJ[4 * wn + bx], off)
                check_conflict_free(addrs)
                for bx in range(2):
                    for by in range(4):
                        for ea in range(2):
                            for eb in range(2):
                                dmma(acc[2 * bx + ea][2 * by + eb], [a[bx][l][ea] for l in range(32)], [b[by][l][eb] for l in range(32)])
        acc_all[warp] = acc
    # second-stage decode (syrk_reduce_kernel)
    out = np.full((BT, BT), np.nan)
    for e in range(BT * BT):
        aidx, tid = e >> 8, e & 255
        warp = tid >> 5
        wm, wn, g, t = warp >> 1, warp & 1, (tid & 31) >> 2, tid & 3
        mt, nt, c = aidx >> 4, (aidx >> 1) & 7, aidx & 1
        i = 32 * wm + 16 * (mt >> 1) + 2 * g + (mt & 1)
        j = 64 * wn + 16 * (nt >> 1) + 4 * t + 2 * c + (nt & 1)
        assert np.isnan(out[i, j])
        out[i, j] = acc_all[warp][mt][nt][tid & 31][c]
    ref = (A[:, :BT] * w[:, None]).T @ A[:, BT:]
    assert np.allclose(out, ref, rtol=1e-13, atol=1e-13)


def test_folded_triangle_enumeration_covers_every_tile_once():
    def slot_to_tile(T, l):
        p, q = divmod(l, T + 1)
        if q < T - p:
            return p, p + q
        p2 = T - 1 - p
        if p2 <= p:
            return None
        return p2, p2 + (q - (T - p))

    for T in (1, 2, 3, 4, 7, 8, 157):
        slots = ((T + 1) // 2) * (T + 1)
        seen = [slot_to_tile(T, l) for l in range(slots)]
        tiles = [s for s in seen if s is not None]
        assert len(tiles) == len(set(tiles)) == T * (T + 1) // 2
        assert all(0 <= ti <= tj < T for ti, tj in tiles)


def _syrk_warp_block(T, u, w):
    """Python restatement of syrk_warp_block (csrc/syrk.cuh): work unit u, warp w -> (cb_a, wm, cb_b, wn, valid, vec, cb0, cb1)."""
    noff = T * (T - 1) // 2
    if u < noff:
        ti, rem = 0, u
        while rem >= T - 1 - ti:
            rem -= T - 1 - ti
            ti += 1
        cb0, cb1 = ti, ti + 1 + rem
        return cb0, w >> 1, cb1, w & 1, True, False, cb0, cb1
    q = u - noff
    first, last = 8 * q, min(8 * q + 7, 6 * T - 1)
    cb0, cb1 = first // 6, last // 6
    idx = first + w
    valid = idx < 6 * T
    t, k = (idx // 6, idx % 6) if valid else (cb0, 0)
    wm = (k >> 1) if k < 4 else (k - 2)
    wn = (k & 1) if k < 4 else 1
    return t, wm, t, wn, valid, valid and k in (0, 2, 4, 5), cb0, cb1


def test_syrk_work_units_cover_upper_triangle_once():
    """Off-diagonal tiles + packed diagonal warp blocks: every (i, j), j >= i, of the K x K result is produced exactly once,
    every 32-column slab has exactly one moment-vector writer, and a unit never needs more than two column blocks."""
    for T in (1, 2, 3, 4, 8, 9):
        K = 128 * T
        units = T * (T - 1) // 2 + (6 * T + 7) // 8
        cover = np.zeros((K, K), dtype=np.int32)
        vec_writers = np.zeros(K, dtype=np.int32)
        for u in range(units):
            for w in range(8):
                cb_a, wm, cb_b, wn, valid, vec, cb0, cb1 = _syrk_warp_block(T, u, w)
                assert cb_a in (cb0, cb1) and cb_b in (cb0, cb1) and cb1 - cb0 >= 0
                if not valid:
                    continue
                i0, j0 = 128 * cb_a + 32 * wm, 128 * cb_b + 64 * wn
                ii, jj = np.meshgrid(np.arange(i0, i0 + 32), np.arange(j0, j0 + 64), indexing="ij")
                keep = jj >= ii
                cover[ii[keep], jj[keep]] += 1
                if vec:
                    vec_writers[i0:i0 + 32] += 1
        assert np.array_equal(np.triu(cover), np.triu(np.ones((K, K), dtype=np.int32))), T
        assert np.all(vec_writers == 1)
    assert 4 * 3 // 2 + (6 * 4 + 7) // 8 == 9  # K = 500: 9 units instead of 10 tiles
