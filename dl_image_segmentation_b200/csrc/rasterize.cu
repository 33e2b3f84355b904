// rasterize.cu — label rasterisation: polygons -> (H, W) uint8 with GDAL's ALL_TOUCHED semantics.
//
// Replaces gdal.RasterizeLayer(mem_ds, [1], layer, options=['ALL_TOUCHED=TRUE'[, 'ATTRIBUTE=...']]) on a background-filled
// Byte raster, create_label_array_for_tile (_descartes_img_chips.py:633-689).  GDAL burns feature after feature, so a pixel
// shared by several polygons keeps the value of the LAST one (the reference's comment, :676-683); here every burn is an
// atomicMax of the feature's index into a per-pixel owner word and one last pass turns owners into values: the same
// result whatever the order the threads run in.  Two burn passes per feature, as in GDAL's alg/llrasterize.cpp:
//   fill  (GDALdllImageFilledPolygon): one thread per (feature, scanline): crossings of the pixel-centre line y + 0.5 with
//         every edge (half-open in y), rounded floor(x + 0.5), sorted, spans between pairs burnt (even-odd rule);
//   lines (GDALdllImageLineAllTouched): one thread per edge, stepping from pixel boundary to pixel boundary.
// All coordinate arithmetic is IEEE double with explicit round-to-nearest operations (no fused multiply-add), in the
// order oracle/rasterize.py states it, so the burnt set is bit-identical with the CPU restatement.
#include "common.cuh"

namespace b2 {
namespace {

__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ void burn(int32_t* owner, int W, int y, int x, int f) { atomicMax(owner + (size_t)y * W + x, f); }

// which feature does job / segment `i` belong to: last f with off[f] <= i
__device__ __forceinline__ int find_owner(const uint32_t* __restrict__ off, int n, uint32_t i) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128)
raster_fill_kernel(const double4* __restrict__ segs, const uint32_t* __restrict__ seg_off, const int32_t* __restrict__ feat_miny,
                   const uint32_t* __restrict__ job_off, int n_features, uint32_t n_jobs, int W, int H, uint32_t max_ints,
                   int32_t* __restrict__ ints_ws, int32_t* __restrict__ owner) {
    const uint32_t job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= n_jobs) return;
    const int f = find_owner(job_off, n_features, job);
    const int y = feat_miny[f] + (int)(job - job_off[f]);
    const double dy = (double)y + 0.5;
    int32_t* ints = ints_ws + (size_t)job * max_ints;
    uint32_t n = 0;
    for (uint32_t s = seg_off[f]; s < seg_off[f + 1]; s++) {
        const double4 e = segs[s];                                   // (x1, y1) -> (x2, y2), ind1 = the earlier vertex
        double dy1 = e.y, dy2 = e.w, dx1, dx2;
        if ((dy1 < dy && dy2 < dy) || (dy1 > dy && dy2 > dy)) continue;
        if (dy1 < dy2) { dx1 = e.x; dx2 = e.z; }
        else if (dy1 > dy2) { dy2 = e.y; dy1 = e.w; dx2 = e.x; dx1 = e.z; }
        else {                                                       // the edge lies on the centre line
            if (e.x > e.z) {
                const int hx1 = (int)floor(dadd(e.z, 0.5)), hx2 = (int)floor(dadd(e.x, 0.5));
                if (!(hx1 > W - 1 || hx2 <= 0))
                    for (int x = max(hx1, 0); x <= min(hx2 - 1, W - 1); x++) burn(owner, W, y, x, f);
            }
            continue;
        }
        if (dy < dy2 && dy >= dy1) {
            const double inter = dadd(ddiv(dmul(dsub(dy, dy1), dsub(dx2, dx1)), dsub(dy2, dy1)), dx1);
            if (n < max_ints) ints[n] = (int)floor(dadd(inter, 0.5));
            n++;
        }
    }
    if (n > max_ints) n = max_ints;                                  // cannot happen: max_ints = the feature's edge count
    for (uint32_t i = 1; i < n; i++) {                               // crossings per scanline are few: insertion sort
        const int32_t v = ints[i];
        uint32_t j = i;
        while (j > 0 && ints[j - 1] > v) { ints[j] = ints[j - 1]; j--; }
        ints[j] = v;
    }
    for (uint32_t i = 0; i + 1 < n; i += 2) {
        const int a = ints[i], b = ints[i + 1];
        if (a <= W - 1 && b > 0)
            for (int x = max(a, 0); x <= min(b - 1, W - 1); x++) burn(owner, W, y, x, f);
    }
}

__global__ void __launch_bounds__(128)
raster_lines_kernel(const double4* __restrict__ segs, const uint32_t* __restrict__ seg_feat, uint32_t n_segs, int W, int H,
                    int32_t* __restrict__ owner) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const double4 e = segs[s];
    const int f = (int)seg_feat[s];
    double x = e.x, y = e.y, xe = e.z, ye = e.w;
    const double dW = (double)W, dH = (double)H;
    if ((y > dH && ye > dH) || (y < 0.0 && ye < 0.0) || (x > dW && xe > dW) || (x < 0.0 && xe < 0.0)) return;
    if (x > xe) { double t = x; x = xe; xe = t; t = y; y = ye; ye = t; }
    if (floor(x) == floor(xe) || fabs(dsub(x, xe)) < 0.01) {         // vertical
        if (ye < y) { const double t = y; y = ye; ye = t; }
        const int ix = (int)floor(xe);
        int iy = (int)floor(y), iye = (int)floor(ye);
        if (ix < 0 || ix >= W) return;
        iy = max(iy, 0);
        iye = min(iye, H - 1);
        for (; iy <= iye; iy++) burn(owner, W, iy, ix, f);
        return;
    }
    if (floor(y) == floor(ye) || fabs(dsub(y, ye)) < 0.01) {         // horizontal
        int ix = (int)floor(x);
        const int iy = (int)floor(y);
        int ixe = (int)floor(xe);
        if (iy < 0 || iy >= H) return;
        ix = max(ix, 0);
        ixe = min(ixe, W - 1);
        for (; ix <= ixe; ix++) burn(owner, W, iy, ix, f);
        return;
    }
    const double slope = ddiv(dsub(ye, y), dsub(xe, x));
    if (xe > dW) { ye = dsub(ye, dmul(dsub(xe, dW), slope)); xe = dW; }
    if (x < 0.0) { y = dadd(y, dmul(dsub(0.0, x), slope)); x = 0.0; }
    if (ye > y) {
        if (y < 0.0) { x = dadd(x, ddiv(dsub(0.0, y), slope)); y = 0.0; }
        if (ye >= dH) xe = dadd(xe, ddiv(dsub(ye, dH), slope));
    } else {
        if (y >= dH) { x = dadd(x, ddiv(dsub(dH, y), slope)); y = dH; }
        if (ye < 0.0) xe = dsub(xe, ddiv(dsub(ye, 0.0), slope));
    }
    while (x >= 0.0 && x < xe) {
        const int ix = (int)floor(x), iy = (int)floor(y);
        if (iy >= 0 && iy < H && ix < W) burn(owner, W, iy, ix, f);
        double step_x = dsub(floor(dadd(x, 1.0)), x);
        double step_y = dmul(step_x, slope);
        if ((int)floor(dadd(y, step_y)) == iy) {
            x = dadd(x, step_x);
            y = dadd(y, step_y);
        } else if (slope < 0) {
            step_y = dsub((double)iy, y);
            if (step_y > -0.000000001) step_y = -0.000000001;
            step_x = ddiv(step_y, slope);
            x = dadd(x, step_x);
            y = dadd(y, step_y);
        } else {
            step_y = dsub((double)(iy + 1), y);
            if (step_y < 0.000000001) step_y = 0.000000001;
            step_x = ddiv(step_y, slope);
            x = dadd(x, step_x);
            y = dadd(y, step_y);
        }
    }
}

__global__ void raster_init_kernel(int32_t* owner, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) owner[i] = -1;
}

__global__ void raster_finish_kernel(const int32_t* __restrict__ owner, const uint8_t* __restrict__ values, uint8_t background,
                                     uint8_t* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int32_t f = owner[i];
        out[i] = f < 0 ? background : values[f];
    }
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" int b2_rasterize_polygons(b2_ctx* ctx, const double* fill_segs, const uint32_t* fill_seg_off, const int32_t* feat_miny,
                                     const uint32_t* job_off, uint32_t n_jobs, const double* line_segs, const uint32_t* line_seg_feat,
                                     uint32_t n_line_segs, const uint8_t* values, int n_features, int width, int height,
                                     int background, uint32_t max_ints, int32_t* owner_ws, int32_t* ints_ws, uint8_t* out,
                                     b2_stream stream) {
    B2_REQUIRE(ctx && out && owner_ws, "b2_rasterize_polygons: NULL argument");
    B2_REQUIRE(width > 0 && height > 0 && (uint64_t)width * height < (1ull << 31), "b2_rasterize_polygons: bad raster size");
    B2_REQUIRE(n_features >= 0 && background >= 0 && background <= 255, "b2_rasterize_polygons: bad argument");
    B2_REQUIRE(n_features == 0 || (fill_seg_off && feat_miny && job_off && values), "b2_rasterize_polygons: NULL feature table");
    B2_REQUIRE(n_jobs == 0 || (fill_segs && ints_ws && max_ints > 0), "b2_rasterize_polygons: NULL fill table");
    B2_REQUIRE(n_line_segs == 0 || (line_segs && line_seg_feat), "b2_rasterize_polygons: NULL line table");
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t npix = (size_t)width * height;
    const unsigned gx = (unsigned)std::min<size_t>((npix + 255) / 256, (size_t)ctx->sm_count * 8);
    raster_init_kernel<<<gx, 256, 0, s>>>(owner_ws, npix);
    ctx->launches++;
    if (n_jobs) {
        raster_fill_kernel<<<(n_jobs + 127) / 128, 128, 0, s>>>(reinterpret_cast<const double4*>(fill_segs), fill_seg_off, feat_miny,
                                                                  job_off, n_features, n_jobs, width, height, max_ints, ints_ws, owner_ws);
        ctx->launches++;
    }
    if (n_line_segs) {
        raster_lines_kernel<<<(n_line_segs + 127) / 128, 128, 0, s>>>(reinterpret_cast<const double4*>(line_segs), line_seg_feat,
                                                                        n_line_segs, width, height, owner_ws);
        ctx->launches++;
    }
    raster_finish_kernel<<<gx, 256, 0, s>>>(owner_ws, values, (uint8_t)background, out, npix);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}
