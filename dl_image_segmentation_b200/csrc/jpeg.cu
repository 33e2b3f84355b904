// jpeg.cu — K1j: baseline JPEG decode for .jpg chips (b2chips.h, "K1j").
// Replaces tf.image.decode_jpeg behind ImageCoder.decode_jpeg (_img_to_tf_threaded.py:36-38,51-56,97-103): libjpeg's
// default pipeline — Huffman decode (ITU-T T.81 Annex F), accurate integer inverse DCT, triangle-filter chroma
// upsampling, fixed-point YCbCr -> RGB — restated for the GPU in three kernels:
//   jpeg_entropy_kernel   one warp per file; the warp builds 10-bit look-ahead tables in shared memory, lane 0 walks the
//                         bit stream (the only serial part) and scatters the non-zero coefficients of each block
//   jpeg_idct_kernel      one thread per 8x8 block: dequantise + column pass + row pass in registers, 8-byte row stores
//   jpeg_colour_kernel    one thread per output pixel: upsample every component at that pixel, convert, write HWC
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"

namespace b2 {
namespace {

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t h_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kLook = 10;  // look-ahead bits

struct HuffTable {              // the DC or AC table of one component, in shared memory
    uint16_t look[1 << kLook];  // (length << 8) | symbol for codes of <= kLook bits, 0 = longer code
    uint32_t lj[17];            // lj[l] = first 16-bit left-justified value beyond the codes of length <= l
    int32_t base[17];           // symbol index of a length-l code = base[l] + code
    uint8_t syms[256];
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t acc;
    int n;          // valid bits in acc (low n bits)
    int marker;     // 0 none, else the marker byte met in the stream (bytes after it read as zero, as libjpeg does)
    int starved;    // zero bytes were fed past a marker / the end of the data

    __device__ __forceinline__ void fill() {
        while (n <= 56) {
            uint32_t b = 0;
            if (marker == 0 && p < end) {
                b = *p;
                if (b == 0xFF) {
                    const uint32_t nb = (p + 1 < end) ? p[1] : 0xD9u;
                    if (nb == 0) {
                        p += 2;
                    } else {
                        marker = (int)nb;
                        b = 0;
                        starved++;
                    }
                } else {
                    p++;
                }
            } else {
                starved++;
            }
            acc = (acc << 8) | b;
            n += 8;
        }
    }
    // make at least 32 bits available: four stream bytes at once when none of them is 0xFF (no stuffing, no marker)
    __device__ __forceinline__ void refill32() {
        if (marker == 0 && p + 8 <= end) {
            const uint32_t* a = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(3));
            const uint32_t lo = __ldg(a), hi = __ldg(a + 1);
            const uint32_t x = __funnelshift_r(lo, hi, 8 * (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3));  // bytes p..p+3, LE
            if ((((~x) - 0x01010101u) & x & 0x80808080u) == 0) {
                acc = (acc << 32) | __byte_perm(x, 0, 0x0123);
                n += 32;
                p += 4;
                return;
            }
        }
        fill();
    }
    // the next 32 bits of the stream, left-aligned (call with n >= 32)
    __device__ __forceinline__ uint32_t top32() const { return (uint32_t)(acc >> (n - 32)); }
};

// code at the top of the left-aligned window w -> (length << 8) | symbol, or 0 for "no such code"
__device__ __forceinline__ uint32_t huff_lookup(uint32_t w, const HuffTable& t) {
    const uint32_t e = t.look[w >> (32 - kLook)];
    if (e) return e;
    // longer codes: canonical codes grow with their length, so the length is a count of comparisons (no branches,
    // the loads are independent of one another)
    const uint32_t t16 = w >> 16;
    int l = kLook + 1;
#pragma unroll
    for (int i = kLook + 1; i < 16; i++) l += t16 >= t.lj[i];
    if (t16 >= t.lj[16]) return 0;
    return ((uint32_t)l << 8) | t.syms[(t.base[l] + (int)(t16 >> (16 - l))) & 255];
}

__device__ __forceinline__ int extend(uint32_t v, int s) { return (s && v < (1u << (s - 1))) ? (int)v - (1 << s) + 1 : (int)v; }

// one warp per file
__global__ void __launch_bounds__(32) jpeg_entropy_kernel(const uint8_t* __restrict__ blob, const b2_jpeg_info* __restrict__ infos,
                                                          const b2_jpeg_job* __restrict__ jobs, int n_jobs,
                                                          int16_t* __restrict__ coef, int32_t* __restrict__ status) {
    __shared__ HuffTable tabs[6];
    __shared__ uint8_t zz[64];
    const int lane = threadIdx.x;
    zz[lane] = c_zigzag[lane];
    zz[lane + 32] = c_zigzag[lane + 32];
    for (int j = blockIdx.x; j < n_jobs; j += gridDim.x) {
        const b2_jpeg_info& fi = infos[j];
        const b2_jpeg_job job = jobs[j];
        __syncwarp();
        // ---- tables: slot c = DC table of component c, slot 3 + c = its AC table
        for (int slot = 0; slot < 6; slot++) {
            const int c = slot % 3, cls = slot / 3;
            if (c >= fi.components) continue;
            const int id = cls ? fi.ta[c] : fi.td[c];
            HuffTable& t = tabs[slot];
            const uint8_t* counts = fi.huff_counts[cls][id];
            const uint8_t* syms = fi.huff_syms[cls][id];
            for (int i = lane; i < (1 << kLook); i += 32) t.look[i] = 0;
            for (int i = lane; i < 256; i += 32) t.syms[i] = syms[i];
            __shared__ int32_t valptr[17], mincode[17];
            if (lane == 0) {
                int code = 0, k = 0;
                for (int l = 1; l <= 16; l++) {
                    valptr[l] = k;
                    mincode[l] = code;
                    t.base[l] = k - code;
                    code += counts[l - 1];
                    k += counts[l - 1];
                    t.lj[l] = (uint32_t)code << (16 - l);
                    code <<= 1;
                }
                valptr[0] = k;  // total number of symbols
            }
            __syncwarp();
            const int total = valptr[0];
            for (int k = lane; k < total; k += 32) {
                int l = 1;
                while (l < 16 && valptr[l + 1] <= k) l++;
                if (l <= kLook) {
                    const int code = mincode[l] + (k - valptr[l]);
                    const int first = code << (kLook - l), cnt = 1 << (kLook - l);
                    const uint16_t e = (uint16_t)((l << 8) | syms[k]);
                    for (int i = 0; i < cnt; i++) t.look[first + i] = e;
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            BitReader br;
            br.p = blob + job.src_off + fi.scan_off;
            br.end = blob + job.src_off + job.src_len;
            br.acc = 0;
            br.n = 0;
            br.marker = 0;
            br.starved = 0;
            int err = 0;
            int pred[3] = {0, 0, 0};
            const int n_mcu = fi.mcus_across * fi.mcus_down;
            uint64_t comp_base[3];
            {
                uint64_t o = job.coef_off;
                for (int c = 0; c < fi.components; c++) {
                    comp_base[c] = o;
                    o += (uint64_t)fi.mcus_across * fi.h[c] * fi.mcus_down * fi.v[c] * 64;
                }
            }
            int next_rst = 0, until_rst = fi.restart_interval;
            for (int m = 0; m < n_mcu && !err; m++) {
                if (fi.restart_interval) {
                    if (until_rst == 0) {
                        if (br.starved * 8 > br.n) {  // the interval used bits beyond its own data
                            err = 2;
                            break;
                        }
                        br.acc = 0;
                        br.n = 0;
                        br.starved = 0;
                        if (br.marker == 0) {  // fill bytes (FF FF ..) or garbage before the marker
                            while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] != 0 && br.p[1] != 0xFF)) br.p++;
                            if (br.p + 1 < br.end) br.marker = br.p[1];
                        }
                        if (br.marker != 0xD0 + next_rst) {
                            err = 2;
                            break;
                        }
                        br.p += 2;
                        br.marker = 0;
                        next_rst = (next_rst + 1) & 7;
                        until_rst = fi.restart_interval;
                        pred[0] = pred[1] = pred[2] = 0;
                    }
                    until_rst--;
                }
                const int my = m / fi.mcus_across, mx = m - my * fi.mcus_across;
                for (int c = 0; c < fi.components && !err; c++) {
                    const HuffTable& dc = tabs[c];
                    const HuffTable& ac = tabs[3 + c];
                    const int bw = fi.mcus_across * fi.h[c];
                    for (int by = 0; by < fi.v[c] && !err; by++)
                        for (int bx = 0; bx < fi.h[c]; bx++) {
                            int16_t* blk = coef + comp_base[c] + ((uint64_t)(my * fi.v[c] + by) * bw + mx * fi.h[c] + bx) * 64;
                            // one refill check, one table load and one shift pair per symbol: code and value bits
                            // (<= 16 + 11) are both taken from the same left-aligned 32-bit window
                            if (br.n < 32) br.refill32();
                            uint32_t w = br.top32();
                            uint32_t e = huff_lookup(w, dc);
                            int l = (int)(e >> 8), s = (int)(e & 0xFFu);
                            if (e == 0 || s > 11) {
                                err = 2;
                                break;
                            }
                            pred[c] += s ? extend((w << l) >> (32 - s), s) : 0;
                            br.n -= l + s;
                            blk[0] = (int16_t)pred[c];
                            int k = 1;
                            while (k < 64) {
                                if (br.n < 32) br.refill32();
                                w = br.top32();
                                e = huff_lookup(w, ac);
                                if (e == 0) {
                                    err = 2;
                                    break;
                                }
                                l = (int)(e >> 8);
                                const int r = (int)(e >> 4) & 15;
                                s = (int)(e & 15u);
                                br.n -= l + s;
                                if (s == 0) {
                                    if (r != 15) break;
                                    k += 16;
                                    continue;
                                }
                                k += r;
                                if (k > 63) {
                                    err = 2;
                                    break;
                                }
                                blk[zz[k]] = (int16_t)extend((w << l) >> (32 - s), s);
                                k++;
                            }
                            if (err) break;
                        }
                }
            }
            // bits consumed beyond the real data = truncated file (tf.image.decode_jpeg fails: try_recover_truncated=False)
            if (!err && br.starved * 8 > br.n) err = 2;
            if (err) atomicMax(&status[job.image], err);
        }
    }
}

#define F_0_298631336 2446
#define F_0_390180644 3196
#define F_0_541196100 4433
#define F_0_765366865 6270
#define F_0_899976223 7373
#define F_1_175875602 9633
#define F_1_501321110 12299
#define F_1_847759065 15137
#define F_1_961570560 16069
#define F_2_053119869 16819
#define F_2_562915447 20995
#define F_3_072711026 25172

// one 8-point pass of the IJG accurate integer inverse DCT (64-bit intermediates as on LP64 hosts)
template <int DESCALE>
__device__ __forceinline__ void idct8(const int in[8], int out[8]) {
    typedef long long L;
    L z2 = in[2], z3 = in[6];
    L z1 = (z2 + z3) * F_0_541196100;
    L tmp2 = z1 + z3 * (-F_1_847759065);
    L tmp3 = z1 + z2 * F_0_765366865;
    z2 = in[0];
    z3 = in[4];
    L tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
    const L tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7];
    tmp1 = in[5];
    tmp2 = in[3];
    tmp3 = in[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    L z4 = tmp1 + tmp3;
    const L z5 = (z3 + z4) * F_1_175875602;
    tmp0 *= F_0_298631336;
    tmp1 *= F_2_053119869;
    tmp2 *= F_3_072711026;
    tmp3 *= F_1_501321110;
    z1 *= -F_0_899976223;
    z2 *= -F_2_562915447;
    z3 = z3 * (-F_1_961570560) + z5;
    z4 = z4 * (-F_0_390180644) + z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    const L rnd = 1LL << (DESCALE - 1);
    out[0] = (int)((tmp10 + tmp3 + rnd) >> DESCALE);
    out[7] = (int)((tmp10 - tmp3 + rnd) >> DESCALE);
    out[1] = (int)((tmp11 + tmp2 + rnd) >> DESCALE);
    out[6] = (int)((tmp11 - tmp2 + rnd) >> DESCALE);
    out[2] = (int)((tmp12 + tmp1 + rnd) >> DESCALE);
    out[5] = (int)((tmp12 - tmp1 + rnd) >> DESCALE);
    out[3] = (int)((tmp13 + tmp0 + rnd) >> DESCALE);
    out[4] = (int)((tmp13 - tmp0 + rnd) >> DESCALE);
}

__device__ __forceinline__ uint32_t range_limit(int v) {  // libjpeg's post-IDCT table: 10-bit wrap, +128, clamp
    const int s = ((v & 0x3FF) ^ 512) - 512 + 128;
    return (uint32_t)min(max(s, 0), 255);
}

// grid = (ceil(max blocks / 128), n_jobs)
__global__ void __launch_bounds__(128) jpeg_idct_kernel(const b2_jpeg_info* __restrict__ infos, const b2_jpeg_job* __restrict__ jobs,
                                                        const int16_t* __restrict__ coef, uint8_t* __restrict__ planes) {
    const b2_jpeg_info& fi = infos[blockIdx.y];
    const b2_jpeg_job& job = jobs[blockIdx.y];
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t coef_off = job.coef_off, plane_off = job.plane_off;
    int c = 0, bw = 0;
    for (; c < fi.components; c++) {
        bw = fi.mcus_across * fi.h[c];
        const uint32_t nb = (uint32_t)bw * fi.mcus_down * fi.v[c];
        if (b < nb) break;
        b -= nb;
        coef_off += (uint64_t)nb * 64;
        plane_off += (uint64_t)nb * 64;
    }
    if (c >= fi.components) return;
    const int16_t* src = coef + coef_off + (uint64_t)b * 64;
    const uint16_t* q = fi.qt[fi.tq[c]];
    int ws[64];
#pragma unroll
    for (int col = 0; col < 8; col++) {
        int in[8], out[8];
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = (int)src[r * 8 + col] * (int)q[r * 8 + col];
        idct8<11>(in, out);
#pragma unroll
        for (int r = 0; r < 8; r++) ws[r * 8 + col] = out[r];
    }
    const int by = b / bw, bx = b - by * bw;
    uint8_t* dst = planes + plane_off + ((uint64_t)by * 8 * bw + bx) * 8;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        int out[8];
        idct8<18>(&ws[r * 8], out);
        uint2 v;
        v.x = range_limit(out[0]) | (range_limit(out[1]) << 8) | (range_limit(out[2]) << 16) | (range_limit(out[3]) << 24);
        v.y = range_limit(out[4]) | (range_limit(out[5]) << 8) | (range_limit(out[6]) << 16) | (range_limit(out[7]) << 24);
        *reinterpret_cast<uint2*>(dst + (uint64_t)r * bw * 8) = v;
    }
}

// the value of one component at full-resolution pixel (x, y), as libjpeg's jdsample.c method choice produces it
__device__ __forceinline__ int upsampled(const uint8_t* __restrict__ pl, int pw, int dw, int dh, int hr, int vr, int x, int y) {
    const bool fancy = dw > 2;
    if (hr == 1 && vr == 1) return pl[(uint64_t)y * pw + x];
    if (fancy && hr == 2 && vr == 1) {
        const uint8_t* row = pl + (uint64_t)y * pw;
        const int i = x >> 1;
        if (x == 0) return row[0];
        if (x == 2 * dw - 1) return row[dw - 1];
        return (x & 1) ? (3 * row[i] + row[i + 1] + 2) >> 2 : (3 * row[i] + row[i - 1] + 1) >> 2;
    }
    if (fancy && hr == 2 && vr == 2) {
        const int r = y >> 1;
        const int far = (y & 1) ? min(r + 1, dh - 1) : max(r - 1, 0);
        const uint8_t* r0 = pl + (uint64_t)r * pw;
        const uint8_t* r1 = pl + (uint64_t)far * pw;
        const int i = x >> 1;
        const int cs = 3 * r0[i] + r1[i];
        if (x == 0) return (cs * 4 + 8) >> 4;
        if (x == 2 * dw - 1) return (cs * 4 + 7) >> 4;
        if (x & 1) return (3 * cs + 3 * r0[i + 1] + r1[i + 1] + 7) >> 4;
        return (3 * cs + 3 * r0[i - 1] + r1[i - 1] + 8) >> 4;
    }
    if (hr == 1 && vr == 2) {  // no width test for this method in libjpeg's jinit_upsampler
        const int r = y >> 1;
        const int far = (y & 1) ? min(r + 1, dh - 1) : max(r - 1, 0);
        return (3 * pl[(uint64_t)r * pw + x] + pl[(uint64_t)far * pw + x] + ((y & 1) ? 2 : 1)) >> 2;
    }
    return pl[(uint64_t)(y / vr) * pw + x / hr];
}

// grid = (ceil(max pixels / 256) capped, n_jobs)
__global__ void __launch_bounds__(256) jpeg_colour_kernel(const b2_jpeg_info* __restrict__ infos, const b2_jpeg_job* __restrict__ jobs,
                                                          const uint8_t* __restrict__ planes, uint8_t* __restrict__ out) {
    const b2_jpeg_info& fi = infos[blockIdx.y];
    const b2_jpeg_job& job = jobs[blockIdx.y];
    const int W = fi.width, H = fi.height, nc = fi.components;
    int hmax = 1, vmax = 1;
    for (int c = 0; c < nc; c++) {
        hmax = max(hmax, fi.h[c]);
        vmax = max(vmax, fi.v[c]);
    }
    const uint8_t* pl[3];
    int pw[3], dw[3], dh[3];
    {
        uint64_t o = job.plane_off;
        for (int c = 0; c < nc; c++) {
            pl[c] = planes + o;
            pw[c] = fi.mcus_across * fi.h[c] * 8;
            o += (uint64_t)pw[c] * fi.mcus_down * fi.v[c] * 8;
            dw[c] = (W * fi.h[c] + hmax - 1) / hmax;
            dh[c] = (H * fi.v[c] + vmax - 1) / vmax;
        }
    }
    uint8_t* dst = out + job.out_off;
    const uint32_t n_px = (uint32_t)W * H;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        int v[3];
        for (int c = 0; c < nc; c++) v[c] = upsampled(pl[c], pw[c], dw[c], dh[c], hmax / fi.h[c], vmax / fi.v[c], x, y);
        if (nc == 1) {
            dst[i] = (uint8_t)v[0];
        } else {
            int r = v[0], g = v[1], b = v[2];
            if (fi.ycc) {
                const int cb = v[1] - 128, cr = v[2] - 128;
                r = v[0] + ((91881 * cr + 32768) >> 16);
                b = v[0] + ((116130 * cb + 32768) >> 16);
                g = v[0] + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
                r = min(max(r, 0), 255);
                g = min(max(g, 0), 255);
                b = min(max(b, 0), 255);
            }
            dst[(uint64_t)i * 3 + 0] = (uint8_t)r;
            dst[(uint64_t)i * 3 + 1] = (uint8_t)g;
            dst[(uint64_t)i * 3 + 2] = (uint8_t)b;
        }
    }
}

int probe_fail(int status, const char* msg) {
    set_error(std::string("b2_jpeg_probe: ") + msg);
    return status;
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" int b2_jpeg_probe(const uint8_t* blob, uint64_t size, b2_jpeg_info* info) {
    if (!blob || !info) return probe_fail(1, "NULL argument");
    memset(info, 0, sizeof(*info));
    if (size < 4 || blob[0] != 0xFF || blob[1] != 0xD8) return probe_fail(1, "not a JPEG");
    uint64_t p = 2;
    bool have_q[4] = {false, false, false, false}, have_h[2][4] = {{false, false, false, false}, {false, false, false, false}};
    bool have_frame = false, jfif = false, have_adobe = false;
    int adobe = 0, ids[3] = {0, 0, 0};
    for (;;) {
        if (p + 4 > size) return probe_fail(1, "truncated before the scan");
        if (blob[p] != 0xFF) return probe_fail(1, "marker expected");
        while (p < size && blob[p] == 0xFF) p++;
        if (p >= size) return probe_fail(1, "truncated before the scan");
        const int m = blob[p++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) return probe_fail(1, "EOI before the scan");
        if (p + 2 > size) return probe_fail(1, "truncated segment");
        const uint64_t ln = ((uint64_t)blob[p] << 8) | blob[p + 1];
        if (ln < 2 || p + ln > size) return probe_fail(1, "truncated segment");
        const uint8_t* seg = blob + p + 2;
        const uint64_t sl = ln - 2;
        p += ln;
        if (m == 0xDB) {
            uint64_t q = 0;
            while (q < sl) {
                const int pq = seg[q] >> 4, tq = seg[q] & 15;
                q++;
                if (tq > 3 || pq > 1) return probe_fail(1, "bad DQT");
                if (q + (pq ? 128 : 64) > sl) return probe_fail(1, "short DQT");
                for (int i = 0; i < 64; i++) {
                    const uint16_t v = pq ? (uint16_t)((seg[q + 2 * i] << 8) | seg[q + 2 * i + 1]) : seg[q + i];
                    info->qt[tq][h_zigzag[i]] = v;
                }
                q += pq ? 128 : 64;
                have_q[tq] = true;
            }
        } else if (m == 0xC4) {
            uint64_t q = 0;
            while (q < sl) {
                if (q + 17 > sl) return probe_fail(1, "short DHT");
                const int tc = seg[q] >> 4, th = seg[q] & 15;
                if (tc > 1 || th > 3) return probe_fail(1, "bad DHT");
                int n = 0, code = 0;
                for (int l = 0; l < 16; l++) {
                    n += seg[q + 1 + l];
                    code += seg[q + 1 + l];
                    if (code > (1 << (l + 1))) return probe_fail(1, "over-subscribed Huffman table");
                    code <<= 1;
                }
                if (n > 256 || q + 17 + n > sl) return probe_fail(1, "bad DHT");
                memcpy(info->huff_counts[tc][th], seg + q + 1, 16);
                memset(info->huff_syms[tc][th], 0, 256);
                memcpy(info->huff_syms[tc][th], seg + q + 17, n);
                have_h[tc][th] = true;
                q += 17 + n;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (sl < 6) return probe_fail(1, "short SOF");
            const int prec = seg[0], nc = seg[5];
            info->height = (seg[1] << 8) | seg[2];
            info->width = (seg[3] << 8) | seg[4];
            if (prec != 8) return probe_fail(3, "12-bit JPEG is out of scope");
            if (nc != 1 && nc != 3) return probe_fail(3, "only 1- and 3-component JPEGs are in scope");
            if (sl < (uint64_t)6 + 3 * nc) return probe_fail(1, "short SOF");
            if (info->height == 0 || info->width == 0) return probe_fail(3, "DNL / empty frame");
            // tf.image.decode_jpeg refuses frames of 2^29 bytes or more ("Image too large", jpeg_mem.cc)
            if ((uint64_t)info->height * info->width * nc >= (1ull << 29)) return probe_fail(3, "image too large");
            info->components = nc;
            for (int c = 0; c < nc; c++) {
                ids[c] = seg[6 + 3 * c];
                info->h[c] = seg[7 + 3 * c] >> 4;
                info->v[c] = seg[7 + 3 * c] & 15;
                info->tq[c] = seg[8 + 3 * c];
                if (info->h[c] < 1 || info->h[c] > 4 || info->v[c] < 1 || info->v[c] > 4 || info->tq[c] > 3)
                    return probe_fail(1, "bad component parameters");
            }
            have_frame = true;
        } else if ((m >= 0xC2 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return probe_fail(3, "progressive / lossless / arithmetic-coded JPEG is out of scope");
        } else if (m == 0xDD) {
            if (sl < 2) return probe_fail(1, "short DRI");
            info->restart_interval = (seg[0] << 8) | seg[1];
        } else if (m == 0xE0 && sl >= 5 && memcmp(seg, "JFIF\0", 5) == 0) {
            jfif = true;
        } else if (m == 0xEE && sl >= 12 && memcmp(seg, "Adobe", 5) == 0) {
            have_adobe = true;
            adobe = seg[11];
        } else if (m == 0xDA) {
            if (!have_frame) return probe_fail(1, "SOS before SOF");
            const int nc = info->components;
            if (sl < 1 || seg[0] != nc) return probe_fail(3, "multi-scan sequential JPEG is out of scope");
            if (sl < (uint64_t)4 + 2 * nc) return probe_fail(1, "short SOS");
            for (int c = 0; c < nc; c++) {
                if (seg[1 + 2 * c] != ids[c]) return probe_fail(3, "multi-scan sequential JPEG is out of scope");
                info->td[c] = seg[2 + 2 * c] >> 4;
                info->ta[c] = seg[2 + 2 * c] & 15;
                if (info->td[c] > 3 || info->ta[c] > 3 || !have_h[0][info->td[c]] || !have_h[1][info->ta[c]] ||
                    !have_q[info->tq[c]])
                    return probe_fail(1, "scan refers to a missing table");
            }
            if (seg[1 + 2 * nc] != 0 || seg[2 + 2 * nc] != 63 || seg[3 + 2 * nc] != 0)
                return probe_fail(1, "bad spectral selection for a sequential scan");
            break;
        }
    }
    const int nc = info->components;
    if (nc == 1) info->h[0] = info->v[0] = 1;  // a single-component scan is never interleaved (T.81 A.2.2)
    int hmax = 1, vmax = 1, blocks = 0;
    for (int c = 0; c < nc; c++) {
        hmax = info->h[c] > hmax ? info->h[c] : hmax;
        vmax = info->v[c] > vmax ? info->v[c] : vmax;
        blocks += info->h[c] * info->v[c];
    }
    if (blocks > 10) return probe_fail(1, "MCU too large");
    for (int c = 0; c < nc; c++)
        if (hmax % info->h[c] || vmax % info->v[c]) return probe_fail(3, "fractional sampling ratios are out of scope");
    info->mcus_across = (info->width + 8 * hmax - 1) / (8 * hmax);
    info->mcus_down = (info->height + 8 * vmax - 1) / (8 * vmax);
    if (nc == 1) info->ycc = 0;
    else if (jfif) info->ycc = 1;
    else if (have_adobe) info->ycc = adobe != 0;
    else info->ycc = !(ids[0] == 82 && ids[1] == 71 && ids[2] == 66);
    info->scan_off = (uint32_t)p;
    return 0;
}

extern "C" int b2_jpeg_sizes(const b2_jpeg_info* info, uint64_t* coef_count, uint64_t* plane_bytes, uint64_t* out_bytes) {
    B2_REQUIRE(info && info->components >= 1 && info->components <= 3, "b2_jpeg_sizes: bad info");
    uint64_t blocks = 0;
    for (int c = 0; c < info->components; c++)
        blocks += (uint64_t)info->mcus_across * info->h[c] * info->mcus_down * info->v[c];
    if (coef_count) *coef_count = blocks * 64;
    if (plane_bytes) *plane_bytes = blocks * 64;
    if (out_bytes) *out_bytes = (uint64_t)info->width * info->height * info->components;
    return 0;
}

extern "C" int b2_jpeg_plan_batch(const uint8_t* const* blobs, const uint64_t* sizes, int n, b2_jpeg_info* infos_out,
                                  int32_t* status_out, b2_jpeg_job* jobs_out, uint8_t* stage_host, uint64_t stage_cap,
                                  int n_threads, b2_jpeg_plan* plan) {
    B2_REQUIRE(n >= 0 && plan && (n == 0 || (blobs && sizes && infos_out && status_out && jobs_out)),
               "b2_jpeg_plan_batch: NULL argument");
    memset(plan, 0, sizeof(*plan));
    if (n == 0) {
        plan->filled = 1;
        return 0;
    }
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 32 ? 32 : nt);
    if (nt > n) nt = n;
    std::vector<b2_jpeg_info> all((size_t)n);
    auto run = [&](auto&& fn) {  // fn(first, last) over [0, n) on nt threads
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(fn, (int)((int64_t)n * t / nt), (int)((int64_t)n * (t + 1) / nt));
        fn(0, (int)((int64_t)n / nt));
        for (auto& x : th) x.join();
    };
    run([&](int a, int b) {
        for (int i = a; i < b; i++) status_out[i] = b2_jpeg_probe(blobs[i], sizes[i], &all[i]);
    });
    uint64_t src = 0, coef = 0, plane = 0, out = 0;
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (status_out[i] != 0) continue;
        uint64_t cc = 0, pb = 0, ob = 0;
        b2_jpeg_sizes(&all[i], &cc, &pb, &ob);
        B2_REQUIRE(sizes[i] <= 0xFFFFFFFFull, "b2_jpeg_plan_batch: file larger than 4 GiB");
        infos_out[m] = all[i];
        b2_jpeg_job& j = jobs_out[m];
        j.src_off = src;
        j.coef_off = coef;
        j.plane_off = plane;
        j.out_off = out;
        j.src_len = (uint32_t)sizes[i];
        j.image = i;
        src = (src + sizes[i] + 15) & ~15ull;
        coef += cc;
        plane = (plane + pb + 15) & ~15ull;
        out = (out + ob + 255) & ~255ull;
        m++;
    }
    plan->stage_bytes = src;
    plan->coef_count = coef;
    plan->plane_bytes = plane;
    plan->out_bytes = out;
    plan->n_jobs = m;
    if (stage_host && stage_cap >= src) {
        const int jobs_n = m;
        n = jobs_n;  // run() partitions [0, n)
        if (n > 0) {
            if (nt > n) nt = n;
            run([&](int a, int b) {
                for (int j = a; j < b; j++) memcpy(stage_host + jobs_out[j].src_off, blobs[jobs_out[j].image], jobs_out[j].src_len);
            });
        }
        plan->filled = 1;
    }
    set_error("");  // per-file probe failures are data (status_out), not an error of this call
    return 0;
}

extern "C" int b2_jpeg_decode(b2_ctx* ctx, const uint8_t* blob_dev, const b2_jpeg_info* infos_dev, const b2_jpeg_info* infos_host,
                              const b2_jpeg_job* jobs_dev, const b2_jpeg_job* jobs_host, int n, int16_t* coef_dev,
                              uint64_t coef_count, uint8_t* planes_dev, uint8_t* out_dev, int32_t* status_dev,
                              b2_stream stream) {
    B2_REQUIRE(ctx && blob_dev && infos_dev && infos_host && jobs_dev && jobs_host && coef_dev && planes_dev && out_dev &&
                   status_dev,
               "b2_jpeg_decode: NULL argument");
    B2_REQUIRE(n >= 0 && n <= 65535, "b2_jpeg_decode: 0..65535 files per call");
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint64_t max_blocks = 0, max_px = 0;
    for (int j = 0; j < n; j++) {
        uint64_t cc = 0, pb = 0, ob = 0;
        if (b2_jpeg_sizes(&infos_host[j], &cc, &pb, &ob)) return 1;
        B2_REQUIRE(jobs_host[j].coef_off + cc <= coef_count, "b2_jpeg_decode: coefficient buffer too small");
        B2_REQUIRE(jobs_host[j].plane_off % 16 == 0, "b2_jpeg_decode: plane_off must be a multiple of 16");
        B2_REQUIRE(infos_host[j].scan_off <= jobs_host[j].src_len, "b2_jpeg_decode: scan_off beyond the file");
        max_blocks = cc / 64 > max_blocks ? cc / 64 : max_blocks;
        const uint64_t px = (uint64_t)infos_host[j].width * infos_host[j].height;
        max_px = px > max_px ? px : max_px;
    }
    B2_CUDA(cudaMemsetAsync(coef_dev, 0, coef_count * sizeof(int16_t), s));
    const int warps = n < ctx->sm_count * 32 ? n : ctx->sm_count * 32;
    jpeg_entropy_kernel<<<warps, 32, 0, s>>>(blob_dev, infos_dev, jobs_dev, n, coef_dev, status_dev);
    B2_CUDA(cudaGetLastError());
    jpeg_idct_kernel<<<dim3((unsigned)((max_blocks + 127) / 128), n), 128, 0, s>>>(infos_dev, jobs_dev, coef_dev, planes_dev);
    B2_CUDA(cudaGetLastError());
    uint64_t gx = (max_px + 255) / 256;
    if (gx > 4096) gx = 4096;
    jpeg_colour_kernel<<<dim3((unsigned)gx, n), 256, 0, s>>>(infos_dev, jobs_dev, planes_dev, out_dev);
    B2_CUDA(cudaGetLastError());
    ctx->launches += 3;
    return 0;
}
