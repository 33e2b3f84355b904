// jpeg_enc.cu — K1je: baseline JPEG encode for the convert_png_to_jpg option (b2chips.h, "K1je").
// Replaces tf.image.encode_jpeg(image, format='', quality=100) behind ImageCoder.png_to_jpeg
// (_img_to_tf_threaded.py:36-38, used at :92-95): libjpeg's compressor with default settings — fixed-point
// RGB -> YCbCr, 2x2 chroma down-sampling (bias 1,2,1,2...), accurate integer forward DCT, quantisation by 8 x q with
// rounding half away from zero, dummy blocks beyond a component's own block grid, the standard Huffman tables.
//   jpeg_fdct_kernel          one thread per 8x8 block in coding order: samples (colour conversion / down-sampling with
//                             libjpeg's edge replication rules) -> forward DCT -> quantise -> zig-zag int16
//   jpeg_block_bits_kernel    one thread per block: length of its Huffman code (DC difference against the previous real
//                             block of its component, run lengths)
//   jpeg_scan_bits_kernel     one CTA per image: exclusive scan -> bit offset of every block
//   jpeg_block_write_kernel   one thread per block: its bits into the unstuffed stream (32-bit atomicOr at the seams)
//   jpeg_stuff_kernel         one CTA per image: 0xFF -> 0xFF 0x00 with a second scan
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
    uint32_t n_blocks, out_b = b;
    if (nc == 1) {
        const int bw = (W + 7) >> 3, bh = (H + 7) >> 3;
        n_blocks = (uint32_t)bw * bh;
        comp = 0;
        by = b / bw;
        bx = b - by * bw;
    } else {
        // threads are numbered plane by plane (all luminance blocks row-major, then Cb, then Cr) so that a warp works on
        // one kind of block and on neighbouring pixels; the block is stored at its place in coding order
        const int mw = (W + 15) >> 4, mh = (H + 15) >> 4;
        const uint32_t n_mcu = (uint32_t)mw * mh;
        n_blocks = n_mcu * 6;
        if (b >= n_blocks) return;
        if (b < 4 * n_mcu) {
            comp = 0;
            by = b / (2 * mw);
            bx = b - by * (2 * mw);
            out_b = ((uint32_t)(by >> 1) * mw + (bx >> 1)) * 6 + (by & 1) * 2 + (bx & 1);
        } else {
            comp = b < 5 * n_mcu ? 1 : 2;
            const uint32_t m = b - (3 + comp) * n_mcu;
            by = m / mw;
            bx = m - by * mw;
            out_b = m * 6 + 3 + comp;
        }
    }
    if (b >= n_blocks) return;
    int16_t* dst = coef + job.coef_off + (uint64_t)out_b * 64;
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

// ---- Huffman coding, block-parallel: bits per block -> exclusive scan per image -> every block writes its own bits
// (32-bit atomicOr at the seams) -> byte stuffing with a second scan.  Bit-identical to the serial coder by construction.
struct CodeTables {
    uint32_t codes[4][256];  // (code << 5) | length, per symbol: DC lum, AC lum, DC chroma, AC chroma
};

__device__ __forceinline__ void build_codes(CodeTables& t) {
    if (threadIdx.x < 4) {
        const int k4 = threadIdx.x;
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            for (int i = 0; i < c_std_bits[k4][l - 1]; i++) t.codes[k4][c_std_vals[k4][k++]] = ((uint32_t)code++ << 5) | (uint32_t)l;
            code <<= 1;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t blocks_of(const b2_jpeg_enc_job& job) {
    return job.components == 1 ? (uint32_t)((job.width + 7) >> 3) * ((job.height + 7) >> 3)
                               : (uint32_t)((job.width + 15) >> 4) * ((job.height + 15) >> 4) * 6;
}

// DC of the previous real block of the same component in coding order (0 at the start of the scan)
__device__ __forceinline__ int prev_dc(const int16_t* __restrict__ c0, uint32_t b, int nc) {
    if (nc == 1) return b ? c0[(uint64_t)(b - 1) * 64] : 0;
    const uint32_t j = b % 6;
    if (j >= 4) return b >= 6 ? c0[(uint64_t)(b - 6) * 64] : 0;
    for (int64_t q = (int64_t)b - 1; q >= 0; q--) {  // dummies sit at the right / bottom edge: a handful of steps at most
        if (q % 6 >= 4) continue;
        const int v = c0[(uint64_t)q * 64];
        if (v != kDummy) return v;
    }
    return 0;
}

template <class Sink>
__device__ __forceinline__ void code_block(const int16_t* __restrict__ blk, int pred, const uint32_t* dc, const uint32_t* ac, Sink& sink) {
    if (blk[0] == kDummy) {  // the DC of the block before it, no AC
        sink.put(dc[0] >> 5, dc[0] & 31);
        sink.put(ac[0] >> 5, ac[0] & 31);
        return;
    }
    const int diff = (int)blk[0] - pred;
    int nb = 32 - __clz(abs(diff));
    sink.put(dc[nb] >> 5, dc[nb] & 31);
    if (nb) sink.put((uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1u), nb);
    int r = 0;
#pragma unroll 1
    for (int k = 1; k < 64; k++) {
        const int v = blk[k];
        if (v == 0) {
            r++;
            continue;
        }
        while (r > 15) {
            sink.put(ac[0xF0] >> 5, ac[0xF0] & 31);
            r -= 16;
        }
        nb = 32 - __clz(abs(v));
        const uint32_t e = ac[(r << 4) | nb];
        sink.put(e >> 5, e & 31);
        sink.put((uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u), nb);
        r = 0;
    }
    if (r) sink.put(ac[0] >> 5, ac[0] & 31);
}

struct CountSink {
    uint32_t bits;
    __device__ __forceinline__ void put(uint32_t, int len) { bits += (uint32_t)len; }
};

struct WordSink {  // big-endian bit stream, written as 32-bit words; the words at both ends are shared with neighbours
    uint32_t* w;   // next word
    uint64_t acc;
    int n;         // bits in acc, including the leading zeros that stand for the neighbour's bits
    __device__ __forceinline__ void put(uint32_t code, int len) {
        acc = (acc << len) | code;
        n += len;
        if (n >= 32) {
            n -= 32;
            atomicOr(w++, __byte_perm((uint32_t)(acc >> n), 0, 0x0123));
        }
    }
    __device__ __forceinline__ void flush() {
        if (n) atomicOr(w, __byte_perm((uint32_t)(acc << (32 - n)), 0, 0x0123));
    }
};

// grid = (ceil(max blocks / 128), n images); bits[] is indexed by the global block number coef_off / 64 + b
__global__ void __launch_bounds__(128) jpeg_block_bits_kernel(const b2_jpeg_enc_job* __restrict__ jobs, const int16_t* __restrict__ coef,
                                                              uint64_t* __restrict__ bits) {
    __shared__ CodeTables t;
    build_codes(t);
    const b2_jpeg_enc_job job = jobs[blockIdx.y];
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= blocks_of(job)) return;
    const int16_t* c0 = coef + job.coef_off;
    const int comp = job.components == 1 ? 0 : (b % 6 < 4 ? 0 : (int)(b % 6) - 3);
    CountSink sink{0};
    code_block(c0 + (uint64_t)b * 64, prev_dc(c0, b, job.components), t.codes[comp ? 2 : 0], t.codes[comp ? 3 : 1], sink);
    bits[job.coef_off / 64 + b] = sink.bits;
}

// one CTA per image: bits[] -> exclusive prefix (bit offset of every block); total[j] = bits of the whole scan
__global__ void __launch_bounds__(256) jpeg_scan_bits_kernel(const b2_jpeg_enc_job* __restrict__ jobs, uint64_t* __restrict__ bits,
                                                             uint64_t* __restrict__ total) {
    __shared__ uint64_t warp_sum[8];
    __shared__ uint64_t carry;
    const b2_jpeg_enc_job job = jobs[blockIdx.x];
    const uint32_t nb = blocks_of(job);
    uint64_t* v = bits + job.coef_off / 64;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t x = i < nb ? v[i] : 0;
        uint64_t inc = x;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint64_t before = carry;
        for (int wi = 0; wi < (int)(threadIdx.x >> 5); wi++) before += warp_sum[wi];
        if (i < nb) v[i] = before + inc - x;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[blockIdx.x] = carry;
}

// grid = (ceil(max blocks / 128), n images): every block writes its bits at its offset into the (zeroed) unstuffed stream
__global__ void __launch_bounds__(128) jpeg_block_write_kernel(const b2_jpeg_enc_job* __restrict__ jobs, const int16_t* __restrict__ coef,
                                                               const uint64_t* __restrict__ bits, const uint64_t* __restrict__ total,
                                                               uint8_t* __restrict__ raw) {
    __shared__ CodeTables t;
    build_codes(t);
    const b2_jpeg_enc_job job = jobs[blockIdx.y];
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x, nb = blocks_of(job);
    if (b >= nb) return;
    const int16_t* c0 = coef + job.coef_off;
    const int comp = job.components == 1 ? 0 : (b % 6 < 4 ? 0 : (int)(b % 6) - 3);
    const uint64_t off = bits[job.coef_off / 64 + b];
    WordSink sink;
    sink.w = reinterpret_cast<uint32_t*>(raw + ((job.out_off >> 1) & ~7ull)) + (off >> 5);
    sink.acc = 0;
    sink.n = (int)(off & 31);
    code_block(c0 + (uint64_t)b * 64, prev_dc(c0, b, job.components), t.codes[comp ? 2 : 0], t.codes[comp ? 3 : 1], sink);
    if (b == nb - 1) {  // pad the last byte of the scan with 1-bits
        const int pad = (int)((8 - (total[blockIdx.y] & 7)) & 7);
        if (pad) sink.put((1u << pad) - 1u, pad);
    }
    sink.flush();
}

// one CTA per image: insert a zero byte after every 0xFF
__global__ void __launch_bounds__(256) jpeg_stuff_kernel(const b2_jpeg_enc_job* __restrict__ jobs, const uint64_t* __restrict__ total,
                                                         const uint8_t* __restrict__ raw, uint8_t* __restrict__ out,
                                                         uint32_t* __restrict__ out_len) {
    __shared__ uint32_t warp_sum[8];
    __shared__ uint64_t carry;
    const b2_jpeg_enc_job job = jobs[blockIdx.x];
    const uint64_t nbytes = (total[blockIdx.x] + 7) >> 3;
    const uint8_t* src = raw + ((job.out_off >> 1) & ~7ull);
    uint8_t* dst = out + job.out_off;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nbytes; base += 256 * 16) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * 16;
        uint8_t v[16];
        uint32_t ff = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            v[k] = i0 + k < nbytes ? src[i0 + k] : 0;
            ff += v[k] == 0xFF;
        }
        uint32_t inc = ff;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
        __syncthreads();
        uint64_t before = carry;
        for (int wi = 0; wi < (int)(threadIdx.x >> 5); wi++) before += warp_sum[wi];
        uint64_t o = i0 + before + inc - ff;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (i0 + k < nbytes) {
                if (o + 2 <= job.out_cap) {
                    dst[o] = v[k];
                    if (v[k] == 0xFF) dst[o + 1] = 0;
                }
                o += v[k] == 0xFF ? 2 : 1;
            }
        __syncthreads();
        if (threadIdx.x == 255) carry = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint64_t len = nbytes + carry;
        out_len[blockIdx.x] = len <= job.out_cap ? (uint32_t)len : 0xFFFFFFFFu;
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
    const dim3 per_block((unsigned)((max_blocks + 127) / 128), n);
    jpeg_fdct_kernel<<<per_block, 128, 0, s>>>(pixels_dev, jobs_dev, qt, coef_dev);
    B2_CUDA(cudaGetLastError());
    // workspace: bit count / offset per block, total per image, the unstuffed stream (half the scan bound per image)
    uint64_t out_end = 0;
    for (int j = 0; j < n; j++) {
        const uint64_t e = jobs_host[j].out_off + jobs_host[j].out_cap;
        B2_REQUIRE(jobs_host[j].out_off % 16 == 0, "b2_jpeg_encode_scan: out_off must be a multiple of 16");
        out_end = e > out_end ? e : out_end;
    }
    const uint64_t n_blocks_all = coef_count / 64;
    const uint64_t bits_bytes = (n_blocks_all + (uint64_t)n) * sizeof(uint64_t);
    const uint64_t raw_bytes = ((out_end / 2 + 64) + 15) & ~15ull;
    WsLock ws_lock(ctx);
    if (ws_reserve(ctx, bits_bytes + raw_bytes, s)) return 1;
    uint64_t* bits = static_cast<uint64_t*>(ctx->ws);
    uint64_t* total = bits + n_blocks_all;
    uint8_t* raw = static_cast<uint8_t*>(ctx->ws) + bits_bytes;
    B2_CUDA(cudaMemsetAsync(raw, 0, raw_bytes, s));
    jpeg_block_bits_kernel<<<per_block, 128, 0, s>>>(jobs_dev, coef_dev, bits);
    B2_CUDA(cudaGetLastError());
    jpeg_scan_bits_kernel<<<n, 256, 0, s>>>(jobs_dev, bits, total);
    B2_CUDA(cudaGetLastError());
    jpeg_block_write_kernel<<<per_block, 128, 0, s>>>(jobs_dev, coef_dev, bits, total, raw);
    B2_CUDA(cudaGetLastError());
    jpeg_stuff_kernel<<<n, 256, 0, s>>>(jobs_dev, total, raw, out_dev, out_len_dev);
    B2_CUDA(cudaGetLastError());
    ctx->launches += 5;
    return 0;
}
