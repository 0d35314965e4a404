// Host build of csrc/raster.h for tests/test_raster_cpu.py: walks the tile sequence of every (rank, CTA pair) exactly as
// the device threads do and counts the visits of each (row block, column block).
#include "raster.h"
#include <cstring>

template <int GBU>
static long long walk(int groups, int gj_log2, int nparts, int npairs, int rows_b, int cols_b, int *visits)
{
    long long steps = 0;
    for (int part = 0; part < nparts; ++part)
        for (int pair = 0; pair < npairs; ++pair) {
            vm::TileIter<GBU> ti;
            for (ti.init(pair, npairs, groups, gj_log2, part, nparts); ti.valid(); ti.next()) {
                ++steps;
                const int bi = ti.bi(), bj = ti.bj();
                if (bi < rows_b && bj < cols_b) ++visits[(long long)bi * cols_b + bj];
            }
        }
    return steps;
}

// gbu = row blocks per column block (1: 256-row blocks, 2: 128-row blocks); visits [rows_b][cols_b], zeroed here
extern "C" long long raster_walk(int gbu, int groups, int gj_log2, int nparts, int npairs, int rows_b, int cols_b, int *visits)
{
    std::memset(visits, 0, sizeof(int) * (size_t)rows_b * cols_b);
    return gbu == 1 ? walk<1>(groups, gj_log2, nparts, npairs, rows_b, cols_b, visits)
                    : walk<2>(groups, gj_log2, nparts, npairs, rows_b, cols_b, visits);
}
