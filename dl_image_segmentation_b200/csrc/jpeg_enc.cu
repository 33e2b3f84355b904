// jpeg_enc.cu — K1je: baseline JPEG encode for the convert_png_to_jpg option (b2chips.h, "K1je").
// Replaces tf.image.encode_jpeg(image, format='', quality=100) behind ImageCoder.png_to_jpeg
// (_img_to_tf_threaded.py:36-38, used at :92-95): libjpeg's compressor with default settings — fixed-point
// RGB -> YCbCr, 2x2 chroma down-sampling (bias 1,2,1,2...), accurate integer forward DCT, quantisation by 8 x q with
// rounding half away from zero, dummy blocks beyond a component's own block grid, the standard Huffman tables.
//   jpeg_fdct_kernel          one thread per 8x8 block in coding order: samples (colour conversion / down-sampling with
//                             libjpeg's edge replication rules) -> forward DCT -> quantise -> zig-zag int16
//   jpeg_huff_encode_kernel   one warp per image; lane 0 codes the blocks in order (DC prediction, run lengths, byte
//                             stuffing) into the scan bytes
// The header (SOI .. SOS) is assembled on the host by b2_jpeg_header.
#include <string.h>

#include "common.cuh"

namespace b2 {
namespace {

// ITU-T T.81 Annex K.3.3 tables, in the order libjpeg writes them: DC luminance, AC luminance, DC chrominance, AC chrominance
const uint8_t h_std_bits[4][16] = {
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125},
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}};
const uint8_t h_std_nvals[4] = {12, 162, 12, 162};
const uint8_t h_std_vals[4][162] = {
    {
        0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},
    {
        1, 2, 3, 0, 4, 17, 5, 18, 33, 49, 65, 6, 19, 81, 97, 7, 34, 113, 20, 50, 129, 145, 161, 8, 35, 66, 177, 193, 21,
        82, 209, 240, 36, 51, 98, 114, 130, 9, 10, 22, 23, 24, 25, 26, 37, 38, 39, 40, 41, 42, 52, 53, 54, 55, 56, 57,
        58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85, 86, 87, 88, 89, 90, 99, 100, 101, 102, 103, 104, 105, 106, 115,
        116, 117, 118, 119, 120, 121, 122, 131, 132, 133, 134, 135, 136, 137, 138, 146, 147, 148, 149, 150, 151, 152,
        153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184, 185, 186, 194, 195,
        196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 225, 226, 227, 228, 229, 230,
        231, 232, 233, 234, 241, 242, 243, 244, 245, 246, 247, 248, 249, 250},
    {
        0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},
    {
        0, 1, 2, 3, 17, 4, 5, 33, 49, 6, 18, 65, 81, 7, 97, 113, 19, 34, 50, 129, 8, 20, 66, 145, 161, 177, 193, 9, 35,
        51, 82, 240, 21, 98, 114, 209, 10, 22, 36, 52, 225, 37, 241, 23, 24, 25, 26, 38, 39, 40, 41, 42, 53, 54, 55, 56,
        57, 58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85, 86, 87, 88, 89, 90, 99, 100, 101, 102, 103, 104, 105, 106,
        115, 116, 117, 118, 119, 120, 121, 122, 130, 131, 132, 133, 134, 135, 136, 137, 138, 146, 147, 148, 149, 150,
        151, 152, 153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184, 185, 186,
        194, 195, 196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 226, 227, 228, 229,
        230, 231, 232, 233, 234, 242, 243, 244, 245, 246, 247, 248, 249, 250}};
const uint8_t h_std_ids[4] = {0x00, 0x10, 0x01, 0x11};
// the same tables for the device
__constant__ uint8_t c_std_bits[4][16] = {
    {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0},
    {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 125},
    {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0},
    {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 119}};
__constant__ uint8_t c_std_nvals[4] = {12, 162, 12, 162};
__constant__ uint8_t c_std_vals[4][162] = {
    {
        0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},
    {
        1, 2, 3, 0, 4, 17, 5, 18, 33, 49, 65, 6, 19, 81, 97, 7, 34, 113, 20, 50, 129, 145, 161, 8, 35, 66, 177, 193, 21,
        82, 209, 240, 36, 51, 98, 114, 130, 9, 10, 22, 23, 24, 25, 26, 37, 38, 39, 40, 41, 42, 52, 53, 54, 55, 56, 57,
        58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85, 86, 87, 88, 89, 90, 99, 100, 101, 102, 103, 104, 105, 106, 115,
        116, 117, 118, 119, 120, 121, 122, 131, 132, 133, 134, 135, 136, 137, 138, 146, 147, 148, 149, 150, 151, 152,
        153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184, 185, 186, 194, 195,
        196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 225, 226, 227, 228, 229, 230,
        231, 232, 233, 234, 241, 242, 243, 244, 245, 246, 247, 248, 249, 250},
    {
        0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11},
    {
        0, 1, 2, 3, 17, 4, 5, 33, 49, 6, 18, 65, 81, 7, 97, 113, 19, 34, 50, 129, 8, 20, 66, 145, 161, 177, 193, 9, 35,
        51, 82, 240, 21, 98, 114, 209, 10, 22, 36, 52, 225, 37, 241, 23, 24, 25, 26, 38, 39, 40, 41, 42, 53, 54, 55, 56,
        57, 58, 67, 68, 69, 70, 71, 72, 73, 74, 83, 84, 85, 86, 87, 88, 89, 90, 99, 100, 101, 102, 103, 104, 105, 106,
        115, 116, 117, 118, 119, 120, 121, 122, 130, 131, 132, 133, 134, 135, 136, 137, 138, 146, 147, 148, 149, 150,
        151, 152, 153, 154, 162, 163, 164, 165, 166, 167, 168, 169, 170, 178, 179, 180, 181, 182, 183, 184, 185, 186,
        194, 195, 196, 197, 198, 199, 200, 201, 202, 210, 211, 212, 213, 214, 215, 216, 217, 218, 226, 227, 228, 229,
        230, 231, 232, 233, 234, 242, 243, 244, 245, 246, 247, 248, 249, 250}};

__constant__ uint8_t c_zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t h_zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                          41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                          30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
// Annex K.1 quantisation tables, natural order
const uint8_t h_std_quant[2][64] = {
    {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
     18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99},
    {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99}};

struct QuantTables {
    uint16_t q[2][64];  // natural order
};

void quant_tables(int quality, QuantTables* qt) {  // jpeg_set_quality(quality, force_baseline = TRUE)
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < 64; i++) {
            long v = ((long)h_std_quant[t][i] * scale + 50) / 100;
            qt->q[t][i] = (uint16_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
        }
}

constexpr int16_t kDummy = 0x7FFF;  // DC slot of a dummy block (no real coefficient reaches it)

__device__ __forceinline__ void fdct8(const int d[8], int out[8], bool first) {
    const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    const int sh = first ? 11 : 15, rnd = 1 << (sh - 1);
    if (first) {
        out[0] = (t10 + t11) * 4;
        out[4] = (t10 - t11) * 4;
    } else {
        out[0] = (t10 + t11 + 2) >> 2;
        out[4] = (t10 - t11 + 2) >> 2;
    }
    int z1 = (t12 + t13) * 4433;
    out[2] = (z1 + t13 * 6270 + rnd) >> sh;
    out[6] = (z1 + t12 * (-15137) + rnd) >> sh;
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 = z3 * (-16069) + z5;
    z4 = z4 * (-3196) + z5;
    out[7] = (a4 + z1 + z3 + rnd) >> sh;
    out[5] = (a5 + z2 + z4 + rnd) >> sh;
    out[3] = (a6 + z2 + z3 + rnd) >> sh;
    out[1] = (a7 + z1 + z4 + rnd) >> sh;
}

// grid = (ceil(max blocks / 128), n images)
__global__ void __launch_bounds__(128) jpeg_fdct_kernel(const uint8_t* __restrict__ pixels, const b2_jpeg_enc_job* __restrict__ jobs,
                                                        QuantTables qt, int16_t* __restrict__ coef) {
    const b2_jpeg_enc_job job = jobs[blockIdx.y];
    const int W = job.width, H = job.height, nc = job.components;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    int comp, bx, by;  // component and block position inside it
    uint32_t n_blocks;
    if (nc == 1) {
        const int bw = (W + 7) >> 3, bh = (H + 7) >> 3;
        n_blocks = (uint32_t)bw * bh;
        comp = 0;
        by = b / bw;
        bx = b - by * bw;
    } else {
        const int mw = (W + 15) >> 4, mh = (H + 15) >> 4;
        n_blocks = (uint32_t)mw * mh * 6;
        const uint32_t mcu = b / 6, j = b - mcu * 6;
        const int my = mcu / mw, mx = mcu - my * mw;
        if (j < 4) {
            comp = 0;
            by = 2 * my + (int)(j >> 1);
            bx = 2 * mx + (int)(j & 1);
        } else {
            comp = (int)j - 3;
            by = my;
            bx = mx;
        }
    }
    if (b >= n_blocks) return;
    int16_t* dst = coef + job.coef_off + (uint64_t)b * 64;
    if (nc == 3 && comp == 0 && (by >= ((H + 7) >> 3) || bx >= ((W + 7) >> 3))) {  // beyond the luminance block grid
        dst[0] = kDummy;
        return;
    }
    const uint8_t* src = pixels + job.src_off;
    int ws[64];
#pragma unroll 1  // unrolled, the three sample sources (grey / luminance / chroma) bloat the loop: 2.4 -> 5.7 ms per 2048 images
    for (int r = 0; r < 8; r++) {
        int d[8], o[8];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            int v;
            if (nc == 1) {
                v = src[(uint64_t)min(by * 8 + r, H - 1) * W + min(bx * 8 + c, W - 1)];
            } else if (comp == 0) {
                const uint8_t* p = src + ((uint64_t)min(by * 8 + r, H - 1) * W + min(bx * 8 + c, W - 1)) * 3;
                v = (19595 * p[0] + 38470 * p[1] + 7471 * p[2] + 32768) >> 16;
            } else {
                // bottom edge: rows are first replicated to an even count, then the last DOWN-SAMPLED row is repeated;
                // right edge: replicated at full resolution (jcprepct.c / jcsample.c)
                const int cy = min(by * 8 + r, ((H + 1) >> 1) - 1), cx = bx * 8 + c;
                const int r0 = 2 * cy, r1 = min(2 * cy + 1, H - 1), c0 = min(2 * cx, W - 1), c1 = min(2 * cx + 1, W - 1);
                int s = (cx & 1) ? 2 : 1;
                const int rr[2] = {r0, r1}, cc[2] = {c0, c1};
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        const uint8_t* p = src + ((uint64_t)rr[i] * W + cc[k]) * 3;
                        s += comp == 1 ? (-11059 * p[0] - 21709 * p[1] + 32768 * p[2] + (128 << 16) + 32767) >> 16
                                       : (32768 * p[0] - 27439 * p[1] - 5329 * p[2] + (128 << 16) + 32767) >> 16;
                    }
                v = s >> 2;
            }
            d[c] = v - 128;
        }
        fdct8(d, o, true);
#pragma unroll
        for (int c = 0; c < 8; c++) ws[r * 8 + c] = o[c];
    }
    const uint16_t* q = qt.q[comp ? 1 : 0];
    int res[64];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        int d[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; r++) d[r] = ws[r * 8 + c];
        fdct8(d, o, false);
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int dv = (int)q[r * 8 + c] * 8, a = abs(o[r]);
            const int m = (a + (dv >> 1)) / dv;
            res[r * 8 + c] = o[r] < 0 ? -m : m;
        }
    }
#pragma unroll
    for (int k = 0; k < 64; k++) dst[k] = (int16_t)res[c_zz[k]];
}

struct BitWriter {
    uint8_t* p;
    uint8_t* end;
    uint64_t acc;
    int n;
    bool overflow;
    __device__ __forceinline__ void put(uint32_t code, int len) {
        acc = (acc << len) | (code & ((1u << len) - 1u));
        n += len;
        while (n >= 8) {
            const uint8_t b = (uint8_t)(acc >> (n - 8));
            n -= 8;
            if (p + 2 > end) {
                overflow = true;
                continue;
            }
            *p++ = b;
            if (b == 0xFF) *p++ = 0;
        }
    }
};

// one warp per image
__global__ void __launch_bounds__(32) jpeg_huff_encode_kernel(const b2_jpeg_enc_job* __restrict__ jobs, int n_jobs,
                                                              const int16_t* __restrict__ coef, uint8_t* __restrict__ out,
                                                              uint32_t* __restrict__ out_len) {
    __shared__ uint32_t codes[4][256];  // (code << 5) | length, per symbol
    const int lane = threadIdx.x;
    if (lane < 4) {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            for (int i = 0; i < c_std_bits[lane][l - 1]; i++) codes[lane][c_std_vals[lane][k++]] = ((uint32_t)code++ << 5) | (uint32_t)l;
            code <<= 1;
        }
    }
    __syncwarp();
    if (lane != 0) return;
    for (int j = blockIdx.x; j < n_jobs; j += gridDim.x) {
        const b2_jpeg_enc_job job = jobs[j];
        const int nc = job.components;
        const uint32_t n_blocks = nc == 1 ? (uint32_t)((job.width + 7) >> 3) * ((job.height + 7) >> 3)
                                          : (uint32_t)((job.width + 15) >> 4) * ((job.height + 15) >> 4) * 6;
        BitWriter bw;
        bw.p = out + job.out_off;
        bw.end = bw.p + job.out_cap;
        bw.acc = 0;
        bw.n = 0;
        bw.overflow = false;
        int pred[3] = {0, 0, 0};
        const int16_t* blk = coef + job.coef_off;
        for (uint32_t b = 0; b < n_blocks; b++, blk += 64) {
            const int j6 = nc == 1 ? 0 : (int)(b % 6);
            const int comp = j6 < 4 ? 0 : j6 - 3;
            const uint32_t* dc = codes[comp ? 2 : 0];
            const uint32_t* ac = codes[comp ? 3 : 1];
            if (blk[0] == kDummy) {  // DC of the block before it, no AC
                bw.put(dc[0] >> 5, dc[0] & 31);
                bw.put(ac[0] >> 5, ac[0] & 31);
                continue;
            }
            int v = blk[0];
            int diff = v - (comp == 0 ? pred[0] : comp == 1 ? pred[1] : pred[2]);
            if (comp == 0) pred[0] = v; else if (comp == 1) pred[1] = v; else pred[2] = v;
            int nb = 32 - __clz(abs(diff));
            bw.put(dc[nb] >> 5, dc[nb] & 31);
            if (nb) bw.put((uint32_t)(diff < 0 ? diff - 1 : diff), nb);
            int r = 0;
            for (int k = 1; k < 64; k++) {
                v = blk[k];
                if (v == 0) {
                    r++;
                    continue;
                }
                while (r > 15) {
                    bw.put(ac[0xF0] >> 5, ac[0xF0] & 31);
                    r -= 16;
                }
                nb = 32 - __clz(abs(v));
                const uint32_t e = ac[(r << 4) | nb];
                bw.put(e >> 5, e & 31);
                bw.put((uint32_t)(v < 0 ? v - 1 : v), nb);
                r = 0;
            }
            if (r) bw.put(ac[0] >> 5, ac[0] & 31);
        }
        if (bw.n) bw.put((1u << (8 - bw.n)) - 1u, 8 - bw.n);  // pad the last byte with 1-bits
        out_len[j] = bw.overflow ? 0xFFFFFFFFu : (uint32_t)(bw.p - (out + job.out_off));
    }
}

}  // namespace
}  // namespace b2

using namespace b2;

extern "C" int b2_jpeg_header(int height, int width, int components, int quality, int density_unit, int x_density,
                              int y_density, uint8_t* out, uint64_t cap, uint64_t* len) {
    B2_REQUIRE(out && len, "b2_jpeg_header: NULL argument");
    B2_REQUIRE((components == 1 || components == 3) && height >= 1 && height <= 65535 && width >= 1 && width <= 65535,
               "b2_jpeg_header: 1 or 3 components, 1..65535 pixels a side");
    QuantTables qt;
    quant_tables(quality, &qt);
    uint8_t buf[1024];
    size_t n = 0;
    auto put = [&](int v) { buf[n++] = (uint8_t)v; };
    auto put16 = [&](int v) { put(v >> 8); put(v & 255); };
    put(0xFF); put(0xD8);
    put(0xFF); put(0xE0); put16(16);
    for (const char* s = "JFIF"; *s; s++) put(*s);
    put(0); put(1); put(1); put(density_unit); put16(x_density); put16(y_density); put(0); put(0);
    for (int t = 0; t < (components == 3 ? 2 : 1); t++) {
        put(0xFF); put(0xDB); put16(67); put(t);
        for (int k = 0; k < 64; k++) put(qt.q[t][h_zz[k]]);
    }
    put(0xFF); put(0xC0); put16(8 + 3 * components); put(8); put16(height); put16(width); put(components);
    if (components == 3) {
        put(1); put(0x22); put(0); put(2); put(0x11); put(1); put(3); put(0x11); put(1);
    } else {
        put(1); put(0x11); put(0);
    }
    for (int t = 0; t < (components == 3 ? 4 : 2); t++) {
        put(0xFF); put(0xC4); put16(19 + h_std_nvals[t]); put(h_std_ids[t]);
        for (int i = 0; i < 16; i++) put(h_std_bits[t][i]);
        for (int i = 0; i < h_std_nvals[t]; i++) put(h_std_vals[t][i]);
    }
    put(0xFF); put(0xDA); put16(6 + 2 * components); put(components);
    if (components == 3) {
        put(1); put(0x00); put(2); put(0x11); put(3); put(0x11);
    } else {
        put(1); put(0x00);
    }
    put(0); put(63); put(0);
    *len = n;
    B2_REQUIRE(cap >= n, "b2_jpeg_header: buffer too small");
    memcpy(out, buf, n);
    return 0;
}

extern "C" int b2_jpeg_encode_sizes(int height, int width, int components, uint64_t* coef_count, uint64_t* scan_cap) {
    B2_REQUIRE((components == 1 || components == 3) && height >= 1 && height <= 65535 && width >= 1 && width <= 65535,
               "b2_jpeg_encode_sizes: 1 or 3 components, 1..65535 pixels a side");
    const uint64_t blocks = components == 1 ? (uint64_t)((width + 7) >> 3) * ((height + 7) >> 3)
                                            : (uint64_t)((width + 15) >> 4) * ((height + 15) >> 4) * 6;
    if (coef_count) *coef_count = blocks * 64;
    if (scan_cap) *scan_cap = blocks * 420 + 16;  // 20 + 63 * 26 bits per block at worst, every byte stuffed
    return 0;
}

extern "C" int b2_jpeg_encode_scan(b2_ctx* ctx, const uint8_t* pixels_dev, const b2_jpeg_enc_job* jobs_dev,
                                   const b2_jpeg_enc_job* jobs_host, int n, int quality, int16_t* coef_dev, uint64_t coef_count,
                                   uint8_t* out_dev, uint32_t* out_len_dev, b2_stream stream) {
    B2_REQUIRE(ctx && pixels_dev && jobs_dev && jobs_host && coef_dev && out_dev && out_len_dev, "b2_jpeg_encode_scan: NULL argument");
    B2_REQUIRE(n >= 0 && n <= 65535, "b2_jpeg_encode_scan: 0..65535 images per call");
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint64_t max_blocks = 0;
    for (int j = 0; j < n; j++) {
        uint64_t cc = 0, cap = 0;
        if (b2_jpeg_encode_sizes(jobs_host[j].height, jobs_host[j].width, jobs_host[j].components, &cc, &cap)) return 1;
        B2_REQUIRE(jobs_host[j].coef_off + cc <= coef_count, "b2_jpeg_encode_scan: coefficient buffer too small");
        max_blocks = cc / 64 > max_blocks ? cc / 64 : max_blocks;
    }
    QuantTables qt;
    quant_tables(quality, &qt);
    jpeg_fdct_kernel<<<dim3((unsigned)((max_blocks + 127) / 128), n), 128, 0, s>>>(pixels_dev, jobs_dev, qt, coef_dev);
    B2_CUDA(cudaGetLastError());
    const int warps = n < ctx->sm_count * 32 ? n : ctx->sm_count * 32;
    jpeg_huff_encode_kernel<<<warps, 32, 0, s>>>(jobs_dev, n, coef_dev, out_dev, out_len_dev);
    B2_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return 0;
}
