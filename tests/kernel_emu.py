"""Host emulation of the two sm_100a kernel families over the work-item tables the library builds
(dtfft_b200/csrc/kernels.cu: transpose_tiles_kernel, rows_copy_kernel; tables: kernel_object.cu:
rebuild_tables).  The emulator follows the device code step for step -- peer interleaving
(item = i * shuffle mod total), binary search of the block, multiply-high division of the item index,
tile bounds -- so that a CPU box can check that every table makes the kernels move exactly the
elements the oracle says, each exactly once.  Test infrastructure only."""
import numpy as np

IN_OFF, OUT_OFF, IS1, IS2, OS0, OS1, OS2, ITEM_BEGIN, SHUFFLE, N0, N1, N2, TILES0, TILES1, D0MUL, D0SHR, D1MUL, D1SHR, DEST, BSHIFT = range(20)


def _fast_div(n, mul, shr):
    """blocks.h FastDiv / kernels.cu fast_div: __umulhi(n, mul) >> shr, or n when mul == 0."""
    if mul == 0:
        return n
    return ((n * mul) >> 32) >> shr


def _find_block(blocks, item):
    lo, hi = 0, len(blocks) - 1
    while lo < hi:
        mid = (lo + hi + 1) >> 1
        if blocks[mid][ITEM_BEGIN] <= item:
            lo = mid
        else:
            hi = mid - 1
    return lo


def _decode(d, item):
    local = item - int(d[ITEM_BEGIN])
    assert 0 <= local < 2 ** 31
    q0 = _fast_div(local, int(d[D0MUL]), int(d[D0SHR]))
    t0 = local - q0 * int(d[TILES0])
    q1 = _fast_div(q0, int(d[D1MUL]), int(d[D1SHR]))
    t1 = q0 - q1 * int(d[TILES1])
    return t0, t1, q1


def run_table(table, family, src, dsts, counts=None, grid=None, noshift=False):
    """Emulate one launch.  ``src`` flat array in kernel units (elements for family 'T', access units
    for family 'R'); ``dsts`` = {dest index: flat array} with -1 = the launch's ``out``.
    ``counts`` (same keys) are incremented once per written unit.  ``grid`` = number of CTAs of the
    grid-stride loop (any value must give the same result)."""
    blocks = [tuple(int(v) for v in row) for row in table["blocks"]]
    total = table["total_items"]
    if not blocks or total == 0:
        return
    shuffle = blocks[0][SHUFFLE]
    if family == "T":
        ta, tb = 32 * table["launch"][0], 32 * table["launch"][1]
    else:
        ta, tb = table["launch"][0], table["launch"][1] * table["launch"][2]
    grid = grid or total
    seen = np.zeros(total, np.int32)
    for cta in range(min(grid, total)):
        for it0 in range(cta, total, grid):
            item = (it0 * shuffle) % total if shuffle > 1 else it0
            seen[item] += 1
            d = blocks[_find_block(blocks, item)]
            t0, t1, c = _decode(d, item)
            assert 0 <= t0 < d[TILES0] and 0 <= t1 < d[TILES1] and 0 <= c < d[N2], (item, t0, t1, c)
            a = np.arange(t0 * ta, min((t0 + 1) * ta, d[N0]), dtype=np.int64)[:, None]
            # family T: the tile grid along b starts BSHIFT elements before the box (line-aligned tile boundaries);
            # `noshift` is the launch-time fallback for misaligned base pointers
            sh = d[BSHIFT] if (family == "T" and not noshift) else 0
            b = np.arange(max(t1 * tb - sh, 0), min((t1 + 1) * tb - sh, d[N1]), dtype=np.int64)[None, :]
            if a.size == 0 or b.size == 0:
                continue
            iidx = d[IN_OFF] + c * d[IS2] + a + b * d[IS1]
            if family == "T":
                oidx = d[OUT_OFF] + c * d[OS2] + a * d[OS0] + b
            else:
                oidx = d[OUT_OFF] + c * d[OS2] + a + b * d[OS1]
            dst = dsts[d[DEST]]
            dst[oidx] = src[iidx]
            if counts is not None:
                np.add.at(counts[d[DEST]], oidx.ravel(), 1)
    assert np.all(seen == 1), "the interleaved item order is not a permutation of the item space"
