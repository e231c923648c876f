"""Host-side checks of the shared-memory index maps of the tensor-path K2 backward (csrc/train_mma.cu) and of the banded
up-sampling kernel (csrc/resample.cu): every warp-wide fragment load / store named in the kernels' comments is simulated lane by
lane and must (a) touch each element of its block exactly once and (b) be free of bank conflicts (32 banks of 4 bytes; a
128-bit access is served per quarter warp, a 64-bit one per half warp).  No GPU needed: this is the arithmetic the kernels'
layouts rest on (tile swizzle of the TMA boxes, table pitch 12, dot-row pitch 33, stage pitch 36, the padded staged rows of K4).
"""
import itertools

import numpy as np

LANES = [(lane >> 2, lane & 3) for lane in range(32)]          # (g, tg) of a lane: MMA fragment coordinates


def box_word(row, col):
    """word address of element (row, col) of a 128-byte-swizzled box with 32-float rows: chunk j of row r sits at j ^ (r & 7)"""
    return row * 32 + ((((col >> 2) ^ row) & 7) << 2) + (col & 3)


def conflict_degree(words, width):
    """worst bank multiplicity of one warp instruction; `words` = first word of each lane, `width` words per lane"""
    group = {1: 32, 2: 16, 4: 8}[width]                            # lanes served together
    worst = 1
    for s in range(0, 32, group):
        banks = {}
        for w0 in words[s:s + group]:
            for k in range(width):
                banks.setdefault((w0 + k) % 32, set()).add(w0 + k)
        worst = max(worst, max(len(v) for v in banks.values()))
    return worst


def test_phase_a_feature_fragments_are_whole_chunks_and_conflict_free():
    # lane (g, tg): chunk g of rows 2 tg and 2 tg + 1 of an 8-row block; element j of the chunk is column slot g of column block j
    for cb in range(4):
        for rsel in (0, 1):
            words = []
            seen = set()
            for g, tg in LANES:
                row = cb * 8 + 2 * tg + rsel
                phys = row * 32 + ((g ^ (row & 7)) << 2)
                assert phys == box_word(row, 4 * g)                # the chunk that holds columns 4 g .. 4 g + 3
                words.append(phys)
                for j in range(4):
                    seen.add((row, 4 * g + j))
            assert conflict_degree(words, 4) == 1
            assert seen == {(cb * 8 + 2 * t + rsel, col) for t in range(4) for col in range(32)}


def test_phase_b2_feature_fragments_cover_a_block_and_are_conflict_free():
    # lane (g, tg): chunks 2 tg and 2 tg + 1 of rows g and g + 8 of a 16-row block; column 8 tg + 2 kb (+1) is slot tg (+4) of block kb
    for mb in range(2):
        seen = set()
        for rsel, csel in itertools.product((0, 8), (0, 1)):
            words = []
            for g, tg in LANES:
                row = mb * 16 + g + rsel
                phys = row * 32 + (((2 * tg + csel) ^ g) << 2)
                assert phys == box_word(row, 8 * tg + 4 * csel)
                words.append(phys)
                seen.update((row, 8 * tg + 4 * csel + j) for j in range(4))
            assert conflict_degree(words, 4) == 1
        assert seen == {(mb * 16 + r, col) for r in range(16) for col in range(32)}
    # both operands must agree on the permuted contraction index: slot tg of column block kb <-> column 8 tg + 2 kb
    cols = sorted(8 * tg + 2 * kb + s for tg in range(4) for kb in range(4) for s in (0, 1))
    assert cols == list(range(32))


def test_table_fragments_at_pitch_12_are_conflict_free():
    ld = 12
    # phase A (A operand = table^T): lane (g, tg) reads column g of rows 2 tg, 2 tg + 1 of a block, and columns 8 + g for g < 2
    for rsel in (0, 1):
        assert conflict_degree([(2 * tg + rsel) * ld + g for g, tg in LANES], 1) == 1
        extra = [(2 * tg + rsel) * ld + 8 + g for g, tg in LANES if g < 2]
        assert len(set(w % 32 for w in extra)) == len(extra)
    # phase B1 (B operand = table^T): lane (g, tg) reads columns tg, tg + 4 and 8 + (tg & 1) of row g of an 8-row block
    for col in (0, 4):
        assert conflict_degree([g * ld + tg + col for g, tg in LANES], 1) == 1
    stacked = [g * ld + 8 + (tg & 1) for g, tg in LANES]
    assert conflict_degree(stacked, 1) == 1                        # lanes tg and tg ^ 2 read the same word: a broadcast, not a conflict
    # the weights of a tile use the same pitch: lane (g, tg) reads columns tg (+4) of pixels g, g + 8
    for psel, col in itertools.product((0, 8), (0, 4)):
        assert conflict_degree([(g + psel) * ld + tg + col for g, tg in LANES], 1) == 1


def test_stacked_columns_8_9_carry_the_three_split_products_once_each():
    # contraction slots of the stacked MMA: slot t (t < 4) and slot t + 4; (W part, T part) per slot as in b1_block / load_w
    terms = []
    for tg in range(4):
        col = 8 + (tg & 1)
        terms.append((col, "hi" if tg < 2 else "lo", "hi"))        # slot tg: W_hi T_hi (tg < 2), W_lo T_hi (tg >= 2)
        if tg < 2:
            terms.append((col, "hi", "lo"))                          # slot tg + 4: W_hi T_lo; zero for tg >= 2
    want = {(col, w, t) for col in (8, 9) for w, t in (("hi", "hi"), ("lo", "hi"), ("hi", "lo"))}
    assert set(terms) == want and len(terms) == len(want)


def test_dot_rows_and_stage_rows_are_conflict_free():
    # dots: C fragment of column block j holds (k = g (+8), box columns 8 tg + j and 8 tg + 4 + j) -> red[k][column - o], pitch 33
    for j, half, o in itertools.product(range(4), (0, 4), range(4)):
        assert conflict_degree([g * 33 + 8 * tg + half + j - o + 3 for g, tg in LANES], 1) == 1
    # the pixel-step warps read a dot row with lane = pixel
    assert conflict_degree([5 * 33 + lane for lane in range(32)], 1) == 1
    # gradient block on its way out: C fragment (pixel 16 mb + g (+8), rows 2 tg, 2 tg + 1) -> stage[row][pixel], pitch 36
    for mb, psel, rsel in itertools.product((0, 1), (0, 8), (0, 1)):
        assert conflict_degree([(2 * tg + rsel) * 36 + mb * 16 + g + psel for g, tg in LANES], 1) == 1
    assert conflict_degree([3 * 36 + lane for lane in range(32)], 1) == 1


def test_k3_backward_norm_row_and_feature_term():
    # train_mma_cos.cu: the squared norm of box column 4 g + j ends up in the lanes tg = 0 (after two shuffles over tg) and goes to
    # row 6 of the warp's dot rows; the eight lanes that store a column block's norms hit eight different banks
    for j, o in itertools.product(range(4), range(4)):
        words = [6 * 33 + 4 * g + j - o + 3 for g in range(8)]
        assert len({w % 32 for w in words}) == 8
    cols = sorted(4 * g + j for g in range(8) for j in range(4))
    assert cols == list(range(32))
    # B1 reads the feature of (row, this lane's pixel) back from the swizzled box: consecutive columns of one row
    for row, o in itertools.product(range(8), range(4)):
        assert conflict_degree([box_word(row, min(lane + o, 31)) for lane in range(28)] + [0, 1, 2, 3], 1) <= 2


def test_tile_geometry_28_of_32_columns_cover_every_pixel_once():
    # box column i of class e holds pixel x_nom + i - o_e with o_e = (e hw + x_nom) & 3; tiles advance by 28 pixels
    hw = 51 * 51
    nt = (hw + 27) // 28
    for e in range(4):
        owner = np.full(hw, -1)
        for t in range(nt):
            x0 = 28 * t
            o = (e * hw + x0) & 3
            assert ((e * hw + x0) & ~3) % 4 == 0                   # aligned box origin
            for p in range(28):                                    # nominal pixels of the tile
                col = p + o
                assert 0 <= col < 32
                if x0 + p < hw:
                    assert owner[x0 + p] == -1
                    owner[x0 + p] = t
        assert (owner >= 0).all()


def test_k4_staged_rows_with_one_pad_per_16_pixels():
    # resample.cu: a lane owns a quad of the flattened output, so the lanes of a warp read pixels 4 apart; pixel X of a staged
    # row (two floats: class 0, class 1) sits at float2 position X + (X >> 4)
    worst_dense = worst_padded = 1
    for x0 in range(0, 64):
        for e in range(4):
            xs = [x0 + 4 * lane + e for lane in range(32)]
            worst_dense = max(worst_dense, conflict_degree([2 * x for x in xs], 2))
            worst_padded = max(worst_padded, conflict_degree([2 * (x + (x >> 4)) for x in xs], 2))
    assert worst_dense == 4                                        # the round-1 layout: 4-way conflicts on every sample
    assert worst_padded <= 2                                       # at most one pair of lanes collides (for a few phases of x0)
    pitch = 401 + (401 >> 4) + 1
    assert max(x + (x >> 4) for x in range(401)) < pitch
