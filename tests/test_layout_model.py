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
    """Python restatement of syrk_warp_block (csrc/syrk.cuh): work unit u, warp w ->
    (kind, cb_a, wm, cb_b, wn, valid, vec, cb0, cb1); kind 0 = 32 x 64 block, 1 = 64 x 64 diagonal square."""
    noff = T * (T - 1) // 2
    if u < noff:
        ti, rem = 0, u
        while rem >= T - 1 - ti:
            rem -= T - 1 - ti
            ti += 1
        cb0, cb1 = ti, ti + 1 + rem
        return 0, cb0, w >> 1, cb1, w & 1, True, False, cb0, cb1
    q = u - noff
    two = 2 * q + 1 < T
    cb0 = 2 * q
    cb1 = 2 * q + 1 if two else 2 * q
    ww = w & 3
    sel = ww >> 1
    valid = sel == 0 or two
    t = cb1 if (sel == 1 and two) else cb0
    if w < 4:
        return 1, t, ww & 1, t, ww & 1, valid, False, cb0, cb1
    return 0, t, ww & 1, t, 1, valid, valid, cb0, cb1


def _square_pair(p):
    bi = 0 if p < 4 else (1 if p < 7 else (2 if p < 9 else 3))
    bj = p - (0, 3, 5, 6)[bi]
    return bi, bj


def test_syrk_work_units_cover_upper_triangle_once():
    """Off-diagonal tiles + diagonal units (two diagonal tiles each: four 64 x 64 squares at box granularity + four upper-right
    32 x 64 blocks): every (i, j), j >= i, of the K x K result is produced exactly once by the reduce kernel's decode, every column
    has exactly one moment-vector writer, a unit never needs more than two column blocks, and the two warps of an SM sub-partition
    (w, w + 4) of a diagonal unit carry one square and one block."""
    assert [_square_pair(p) for p in range(10)] == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (1, 2), (1, 3), (2, 2), (2, 3), (3, 3)]
    for T in (1, 2, 3, 4, 8, 9):
        K = 128 * T
        units = T * (T - 1) // 2 + (T + 1) // 2
        cover = np.zeros((K, K), dtype=np.int32)
        vec_writers = np.zeros(K, dtype=np.int32)
        for u in range(units):
            kinds = []
            for w in range(8):
                kind, cb_a, wm, cb_b, wn, valid, vec, cb0, cb1 = _syrk_warp_block(T, u, w)
                kinds.append(kind)
                assert cb_a in (cb0, cb1) and cb_b in (cb0, cb1) and cb1 - cb0 >= 0
                if not valid:
                    continue
                # the reduce kernel's decode of (acc index a, lane (g, t)) -> (i, j)
                g, t = np.meshgrid(np.arange(8), np.arange(4), indexing="ij")
                for a in range(80 if kind == 1 else 64):
                    if kind == 0:
                        mt, nt, c = a >> 4, (a >> 1) & 7, a & 1
                        i = 128 * cb_a + 32 * wm + 16 * (mt >> 1) + 2 * g + (mt & 1)
                        j = 128 * cb_b + 64 * wn + 16 * (nt >> 1) + 4 * t + 2 * c + (nt & 1)
                    else:
                        bi, bj = _square_pair(a >> 3)
                        e1, e2, c = (a >> 2) & 1, (a >> 1) & 1, a & 1
                        base = 128 * cb_a + 64 * wm
                        i = base + 16 * bi + 2 * g + e1
                        j = base + 16 * bj + 4 * t + 2 * c + e2
                    keep = j >= i
                    np.add.at(cover, (i[keep], j[keep]), 1)
                if vec:
                    assert kind == 0
                    a0 = 128 * cb_a + 32 * wm
                    b0 = 128 * cb_b + 64 * wn + 32 * wm
                    vec_writers[a0:a0 + 32] += 1
                    vec_writers[b0:b0 + 32] += 1
            if u >= T * (T - 1) // 2:
                assert all(kinds[w] + kinds[w + 4] == 1 for w in range(4)), "one square and one block per SM sub-partition"
        assert np.array_equal(np.triu(cover), np.triu(np.ones((K, K), dtype=np.int32))), T
        assert np.all(vec_writers == 1)
    assert 4 * 3 // 2 + (4 + 1) // 2 == 8  # K = 500: 6 + 2 units (cost 8.25) instead of 9 packed units / 10 tiles


def _smem_wavefronts(addrs):  # 8-byte accesses of one warp, 32 banks x 4 bytes: greedy packing into conflict-free wavefronts (>= 2)
    waves = 0
    remaining = list(addrs)
    while remaining:
        used, nxt = {}, []
        for a in remaining:
            bank = (a // 8) % 16  # pair of 4-byte banks
            if bank in used and used[bank] != a:
                nxt.append(a)
            else:
                used[bank] = a
        waves += 1
        remaining = nxt
    return waves


def test_staged_epilogue_box_layout():
    """gram_epilogue_staged (csrc/gram.cuh): the warp tile (32 x 64) leaves in two halves through ONE 8 KB per-warp staging buffer that
    holds two 16-column x 32-row boxes in TMA's 128B-swizzled layout (16-byte chunk index XOR (row & 7)). Direct copy: half h = columns
    32h .. 32h+31, box (col >> 4) & 1. Mirrored copy (staged transposed): half ih = i-columns 16 ih .. +15, box jh = j-rows 32 jh .. +31.
    Replays the st.shared address math and checks that (a) each half is a bijection onto the 8 KB buffer, (b) un-swizzling box by box
    gives the warp tile resp. its transpose, (c) a store instruction needs the minimum of 2 (direct) / at most 4 (mirror) shared-memory wavefronts."""
    def perm(g):
        return ((g & 1) << 2) | (g >> 1)

    tile = np.arange(32 * 64, dtype=np.float64).reshape(32, 64)  # value = 64 i + j
    for mirror in (False, True):
        for half in range(2):
            buf = np.full(8192 // 8, np.nan)
            worst = 0
            for mt in range(4):
                for nt in range(8):
                    if (not mirror and nt // 4 != half) or (mirror and mt // 2 != half):
                        continue
                    for c in range(2):
                        addrs = []
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            pg = perm(g)
                            il, jl = 8 * mt + pg, 8 * nt + t + 4 * c
                            if not mirror:
                                a = ((jl >> 4) & 1) * 4096 + il * 128 + (((jl & 15) << 3) ^ (pg << 4))
                            else:
                                a = (nt >> 2) * 4096 + (jl & 31) * 128 + (((il & 15) << 3) ^ ((t + 4 * c) << 4))
                            assert np.isnan(buf[a // 8])
                            buf[a // 8] = tile[il, jl]
                            addrs.append(a)
                        worst = max(worst, _smem_wavefronts(addrs))
            assert not np.isnan(buf).any()
            assert worst <= (4 if mirror else 2)
            # what the TMA unit reads: box k, row r, 16-byte chunk q lives at chunk q ^ (r & 7)
            for box in range(2):
                got = np.empty((32, 16))
                for r in range(32):
                    for col in range(16):
                        a = box * 4096 + r * 128 + ((col << 3) ^ ((r & 7) << 4))
                        got[r, col] = buf[a // 8]
                if not mirror:
                    assert np.array_equal(got, tile[:, 32 * half + 16 * box:32 * half + 16 * box + 16])
                else:
                    assert np.array_equal(got, tile[16 * half:16 * half + 16, 32 * box:32 * box + 32].T)


def test_exp_table_replicas_are_conflict_free():
    """exp2_tab (csrc/gram.cuh): entry j of lane slot s = lane & 15 lives at (16 j + s) * 8 — whatever the 32 indices of a warp are, each
    bank pair is hit exactly twice: 2 wavefronts per lookup, the minimum for 256 bytes (an unreplicated table needs ~8 for random indices)."""
    rng = np.random.default_rng(5)
    for _ in range(200):
        idx = rng.integers(0, 256, size=32)
        assert _smem_wavefronts([int((16 * j + (lane & 15)) * 8) for lane, j in enumerate(idx)]) == 2
    flat = [_smem_wavefronts([int(j) * 8 for j in rng.integers(0, 4096, size=32)]) for _ in range(200)]
    assert np.mean(flat) > 4  # what a 4096-entry unreplicated table costs in this model (ncu measured 6.2 CONFLICT wavefronts per lookup)


def test_rbf_augmented_operands_give_the_exponent():
    """rbf_aug_kernel (csrc/prelude.cuh) + the B operand's shifted last k-step (GramParams::b_shift) + exp2_tab, restated in numpy / C:
    rows scaled by c = sqrt(256 log2 e) / ell after the shift, two augmented columns (u, 1) x (1, u) in the last two places of the last
    16-column block, the B variant of that block stored once more behind the row. The contraction over 16 KT columns is then
    256 log2(e) (-0.5 d2 / ell^2 + ln beta), and 2^(./256) of it is the reference's beta exp(-0.5 d2 / ell^2) (kernels.jl:73-74,
    distances.jl:58-60) to ~1e-15 — for every depth class: columns fit the padding, start a new k-step, sit at its end."""
    import ctypes
    import os
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    src, lib = os.path.join(here, "fast_exp_model.c"), os.path.join(here, "..", "oracle", "_build", "libfastexp_model.so")
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-mfma", "-fPIC", "-shared", "-o", lib, src, "-lm"])
    L = ctypes.CDLL(lib)
    L.exp2_tab_model.restype = ctypes.c_double
    L.exp2_tab_model.argtypes = [ctypes.c_double]
    rng = np.random.default_rng(11)
    Lc = 256.0 * 1.4426950408889634074
    for d in (1, 9, 13, 14, 15, 16, 17, 26, 30, 31, 32, 300):
        n, ell, beta = 37, 0.8 + 0.1 * (d % 5), 0.4 + 0.3 * (d % 4)
        X = rng.standard_normal((n, d)) * 0.6 + 3.0 * rng.standard_normal(d)
        mean = X[:: max(1, (n + 511) // 512)].mean(axis=0)
        KT = (d + 2 + 15) // 16
        ld = 16 * KT + 16
        c, h = np.sqrt(Lc) / abs(ell), 0.5 * Lc * np.log(beta)
        buf = np.full((n, ld), np.nan)
        xs = (X - mean) * c
        u = -0.5 * np.sum(xs * xs, axis=1) + h
        buf[:, :d] = xs
        buf[:, d:16 * KT - 2] = 0.0
        buf[:, 16 * KT - 2], buf[:, 16 * KT - 1] = u, 1.0
        b0 = 16 * (KT - 1)
        buf[:, 16 * KT:16 * KT + 14] = buf[:, b0:b0 + 14]
        buf[:, 16 * KT + 14], buf[:, 16 * KT + 15] = 1.0, u
        assert not np.isnan(buf).any()
        A = buf[:, :16 * KT]
        B = np.concatenate([buf[:, :b0], buf[:, 16 * KT:16 * KT + 16]], axis=1)  # last k-step read 16 columns further right
        Y = A @ B.T
        D2 = np.sum((X[:, None, :] - X[None, :, :]) ** 2, axis=2)
        ref = beta * np.exp(-0.5 * D2 / ell ** 2)
        K = np.array([[L.exp2_tab_model(float(y)) for y in row] for row in Y])
        off = ~np.eye(n, dtype=bool)  # the kernel forces the diagonal to beta exactly
        assert np.max(np.abs(K[off] - ref[off]) / ref[off]) < 5e-13, d
