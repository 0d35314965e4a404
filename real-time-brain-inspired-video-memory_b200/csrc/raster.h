// raster.h -- tile sequence of the all-pairs scorer (pairs.cu), kept in a header of plain integer arithmetic so that the
// SAME code is compiled for the host by tests/test_raster_cpu.py, which checks that every tile of the upper triangle is
// visited exactly once over all ranks and CTA pairs (the multi-GPU dealing has no other CPU-side witness).
#pragma once
#ifdef __CUDACC__
#define VM_RASTER_FN __device__ __forceinline__
#else
#define VM_RASTER_FN inline
#endif

namespace vm {

// Tile sequence of one CTA (pair).  Tiles are numbered group by group; group g holds GBU*J*(g+1) row blocks x J column
// blocks (J = 2^gj_log2 column blocks per group; GBU = 2 for 128-row blocks, 1 for 256-row blocks).  With several ranks
// (part / nparts) a rank owns whole ROW BLOCKS of every group (bi % nparts == part), so it walks the same wide groups as
// a single GPU would -- its column panel stays in its L2 and each of its row blocks is reused across all J columns
// (dealing single tiles cyclically would leave every rank only J / nparts columns per group: measured at 8 GPUs,
// 92 -> 105 ms per pass when J went 8 -> 64 that way).  Inside a rank the tiles go to the CTA pairs cyclically.
// The iterator keeps (group, offset in the rank's part of it) and advances with integer arithmetic only -- it runs on
// the single MMA-issuing thread between tiles.
template <int GBU>
struct TileIter {
    long long u, stride, cur;  // offset inside the rank's tiles of group g; tiles the rank owns in group g
    int g, groups, lg, part, nparts;
    VM_RASTER_FN long long local_tiles(int gg) const
    {
        const long long rows = ((long long)GBU << lg) * (gg + 1);               // row blocks of the group
        return ((rows - part + nparts - 1) / nparts) << lg;                      // those with bi % nparts == part, x J columns
    }
    VM_RASTER_FN void init(long long first, long long stride_, int groups_, int gj_log2, int part_, int nparts_)
    {
        u = first; stride = stride_; g = 0; groups = groups_; lg = gj_log2; part = part_; nparts = nparts_;
        cur = local_tiles(0);
        settle();
    }
    VM_RASTER_FN void settle()
    {
        while (g < groups && u >= cur) { u -= cur; ++g; cur = local_tiles(g); }
    }
    VM_RASTER_FN bool valid() const { return g < groups; }
    VM_RASTER_FN void next() { u += stride; settle(); }
    VM_RASTER_FN int bi() const { return (int)((u >> lg) * nparts + part); }
    VM_RASTER_FN int bj() const { return (int)(((long long)g << lg) + (u & ((1 << lg) - 1))); }
};

}  // namespace vm
