// path_box.cuh -- the path-level test of the Dubins edge verification, one THREAD per path (plan kernel).
//
// `box` = axis-aligned box of every point the sampled path can produce and of the parent point (pp_path_box,
// dubins_device.cuh; only for words with three positive lengths).  The path is FREE -- without a sample being
// generated -- when (i) every cell of the bounds classification grid under the box is "inside" and (ii) no ring
// registered under the box's cells of the obstacle grid has an fp32 box that meets the box: then no sample can fail its
// own bounds test and no sample segment can find a candidate ring in its own walk (its cells and its fp32 box are
// subsets of the path's), so the verify kernel's per-sample loop would return "free" as well.  Applies to boxes of at
// most 32 bounds cells and PP_PATH_BOX_CELLS^2 obstacle cells (short edges: the extend step).
// The first version of this test ran in the verify kernel, one WARP per path, and was a third of that kernel on the
// extend-step workload (ncu r03: pp_path_box_free 33 % of the samples); per thread it costs a 32nd of that, and the
// verify kernel only sees the paths that fail it (a compacted index list).
#pragma once
#include "pp_common.cuh"

__device__ __forceinline__ bool pp_path_box_free_thread(const pp_world_view &w, double x0, double y0, double x1, double y1) {
    if (!((x1 - x0) < 1e300 && (y1 - y0) < 1e300 && x0 <= x1 && y0 <= y1)) return false;  // NaN / inf: no shortcut
    {   // (i) bounds
        const int ix0 = __double2int_rd((x0 - w.bminx) * w.binvx), iy0 = __double2int_rd((y0 - w.bminy) * w.binvy);
        const int ix1 = __double2int_rd((x1 - w.bminx) * w.binvx), iy1 = __double2int_rd((y1 - w.bminy) * w.binvy);
        if (ix0 < 0 || iy0 < 0 || ix1 >= w.bgx || iy1 >= w.bgy) return false;
        const int nx = ix1 - ix0 + 1, ny = iy1 - iy0 + 1;
        if (nx > 32 || ny > 32 || nx * ny > 32) return false;
        for (int iy = iy0; iy <= iy1; ++iy)
            for (int ix = ix0; ix <= ix1; ++ix)
                if (__ldg(w.bcls + (size_t)iy * w.bgx + ix) != 1) return false;
    }
    if (w.n_rings == 0u) return true;
    // (ii) obstacles: the cell range exactly as the per-sample walk computes it
    int cx0 = __double2int_rd((x0 - w.gminx) * w.ginv), cy0 = __double2int_rd((y0 - w.gminy) * w.ginv);
    int cx1 = __double2int_rd((x1 - w.gminx) * w.ginv), cy1 = __double2int_rd((y1 - w.gminy) * w.ginv);
    if (cx1 < 0 || cy1 < 0 || cx0 >= w.gx || cy0 >= w.gy) return true;  // beside the ring grid: nothing is registered there
    cx0 = max(cx0, 0);
    cy0 = max(cy0, 0);
    cx1 = min(cx1, w.gx - 1);
    cy1 = min(cy1, w.gy - 1);
    if (cx1 - cx0 >= PP_PATH_BOX_CELLS || cy1 - cy0 >= PP_PATH_BOX_CELLS) return false;
    const float q0x = __double2float_rd(x0), q0y = __double2float_rd(y0), q1x = __double2float_ru(x1), q1y = __double2float_ru(y1);
    for (int r = cy0; r <= cy1; ++r) {
        const uint32_t *row = w.cell_start + (size_t)r * w.gx;
        const uint32_t k1 = __ldg(row + cx1 + 1);  // a row of cells is one run of the cell-ordered boxes
        for (uint32_t k = __ldg(row + cx0); k < k1; ++k) {
            const float4 bb = __ldg(w.cell_box + k);
            if (!(q1x < bb.x || q0x > bb.z || q1y < bb.y || q0y > bb.w)) return false;
        }
    }
    return true;
}
