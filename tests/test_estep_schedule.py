"""CPU: the work order of the single-launch E-step (csrc/nfh_schedule.h, through the C ABI's introspection call).
Every (phase, individual, tile) item appears exactly once, no posterior item precedes a product item of its own
wave (the no-deadlock argument of the kernel), and posterior items walk a wave in the reverse of product order."""
import ctypes as C
import itertools

import pytest

import ngsf_hmm_b200 as nfh


def order(n_rows, n_tiles, wave_rows, lookahead):
    lib = nfh.load_library()
    f = lib.nfh_estep_schedule_item
    f.restype = C.c_uint64
    f.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32)]
    out = (C.c_uint32 * 3)()
    total = f(n_rows, n_tiles, wave_rows, lookahead, 2 ** 62, out)
    items = []
    for t in range(total):
        f(n_rows, n_tiles, wave_rows, lookahead, t, out)
        items.append((out[0], out[1], out[2]))
    return items


CASES = [(1, 1, 1, 0), (1, 7, 1, 3), (5, 3, 2, 0), (5, 3, 2, 4), (5, 3, 2, 100), (7, 4, 7, 5), (7, 4, 100, 5),
         (9, 2, 3, 1), (10, 5, 4, 6), (4, 6, 1, 2), (13, 3, 5, 7), (6, 1, 2, 1)]


@pytest.mark.parametrize("n_rows,n_tiles,wave_rows,lookahead", CASES)
def test_every_item_once_and_products_first(n_rows, n_tiles, wave_rows, lookahead):
    items = order(n_rows, n_tiles, wave_rows, lookahead)
    assert len(items) == 2 * n_rows * n_tiles
    assert sorted(items) == sorted(itertools.product((0, 1), range(n_rows), range(n_tiles)))
    wr = min(wave_rows, n_rows)
    first_apply = {}
    last_product = {}
    for t, (ph, row, tile) in enumerate(items):
        w = row // wr
        if ph:
            first_apply.setdefault(w, t)
        else:
            last_product[w] = t
    for w in first_apply:
        assert last_product[w] < first_apply[w], (w, last_product[w], first_apply[w])
    # posterior items of a wave come in the exact reverse of its product items
    for w in first_apply:
        p = [(r, tl) for ph, r, tl in items if not ph and r // wr == w]
        a = [(r, tl) for ph, r, tl in items if ph and r // wr == w]
        assert a == p[::-1]


def test_lookahead_items_precede_the_posteriors_of_the_previous_wave():
    items = order(6, 4, 2, 3)                # waves of 8 items
    assert [ph for ph, _, _ in items[:8]] == [0] * 8
    assert [ph for ph, _, _ in items[8:11]] == [0, 0, 0]            # lookahead of wave 1
    assert [ph for ph, _, _ in items[11:21]] == [1, 0] * 5           # then posterior / product alternate
    assert [ph for ph, _, _ in items[21:24]] == [1, 1, 1]            # remaining posteriors of wave 0
