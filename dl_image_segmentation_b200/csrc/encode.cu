// encode.cu — K1w: the GeoTIFF *writer* side of the chip format: tile split + TIFF-LZW encode on the GPU.
//
// Replaces (reference call sites): GDAL GTiff driver with COMPRESS=LZW, TILED=TRUE behind
//     _gdal_dataset_from_geocontext + band.WriteArray      _descartes_img_chips.py:781-797, 804-849
// i.e. the step that turns a composite (K3) and its label raster into the chip pair the translators (K1/K2) read.
//
// TIFF-LZW is the exact mirror of the decoder in codec.cu: MSB-first codes, leading Clear, 9 -> 12 bit "early change"
// widths, Clear when the table reaches 4094 entries, EOI.  Greedy longest-match parsing with an exact dictionary is
// deterministic, so the output is byte-identical with any conforming greedy encoder (tests compare it with the CPU
// fixture encoder and decode it back with libtiff).  Encoding is serial inside a stream (every step depends on the
// dictionary built so far): one warp owns one tile, all lanes run the same scalar walk (no divergence) and cooperate
// on what is parallel: clearing the table and staging the input window.  Parallelism comes from the tiles of a batch.
//
// The walk is one dependent chain per input byte (key -> hash -> probe -> compare), so what it costs is the latency
// of the probe: the dictionary — an open-addressing hash table of 6144 words, (key << 12) | code with
// key = (prefix << 8) | byte, at most 4094 entries so the load stays under 0.67 — sits in SHARED memory (one warp per
// CTA, so every address is "register + immediate").  With the table in global memory (L2) a byte cost ~1000 cycles
// and a 512 KiB tile 275 ms; nine resident warps per SM instead of sixty-four is a good trade for a 5x shorter chain
// (59 ms per tile, ~210 cycles per byte on chips whose noisy low bytes make nearly every step a dictionary miss).
#include "common.cuh"

namespace b2 {

constexpr int kEncSlots = 6144;         // hash slots per stream: load <= 0.67 (4094 entries), ~1.6 probes on average
constexpr int kEncWin = 1024;           // staged input bytes per refill
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

struct EncSmem {
    uint32_t table[kEncSlots];
    uint32_t win[kEncWin / 4];
};

struct BitWriter {
    uint8_t* dst;
    uint32_t cap, o;
    uint64_t acc;
    int nacc;
    bool fail;
    __device__ __forceinline__ void put(uint32_t code, int nb, bool writer) {
        acc = (acc << nb) | code;
        nacc += nb;
        if (nacc >= 32) {               // flush four bytes, big-endian, at a 4-byte aligned position
            const uint32_t w = (uint32_t)(acc >> (nacc - 32));
            if (o + 4 <= cap) {
                if (writer) *reinterpret_cast<uint32_t*>(dst + o) = __byte_perm(w, 0, 0x0123);
            } else {
                fail = true;
            }
            o += 4;
            nacc -= 32;
        }
    }
    __device__ __forceinline__ void finish(bool writer) {
        while (nacc > 0) {
            const int take = nacc >= 8 ? 8 : nacc;
            const uint32_t b = (uint32_t)((acc >> (nacc - take)) << (8 - take)) & 0xFFu;
            if (o < cap) {
                if (writer) dst[o] = (uint8_t)b;
            } else {
                fail = true;
            }
            o++;
            nacc -= take;
        }
    }
};

__global__ void __launch_bounds__(32)
lzw_encode_kernel(const uint8_t* __restrict__ raw, const b2_enc_desc* __restrict__ descs, int n, uint8_t* __restrict__ out,
                  uint32_t* __restrict__ out_len, unsigned int* next_stream) {
    __shared__ __align__(16) EncSmem sm;
    const int lane = threadIdx.x;
    const bool writer = lane == 0;
    for (;;) {                                           // persistent warps draw tiles from a counter
        int si = 0;
        if (lane == 0) si = (int)atomicAdd(next_stream, 1u);
        si = __shfl_sync(0xffffffffu, si, 0);
        if (si >= n) return;
        const b2_enc_desc d = descs[si];
        const uint8_t* src = raw + d.src_off;
        const uint32_t len = d.src_len;
        BitWriter w{out + d.dst_off, d.dst_cap, 0, 0, 0, false};
        enum { CLEAR = 256, EOI = 257, FIRST = 258, LIMIT = 4094 };
        int nbits = 9, next = FIRST;
        uint32_t cur = 0;
        auto clear_table = [&]() {
            __syncwarp();
            for (int k = lane; k < kEncSlots / 4; k += 32) reinterpret_cast<uint4*>(sm.table)[k] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
            __syncwarp();
        };
        // stage input bytes [lo, lo + kEncWin) (lo a multiple of 16) with coalesced 16-byte loads, zero past the end
        auto stage = [&](uint32_t lo) {
            __syncwarp();
            for (int k = lane; k < kEncWin / 16; k += 32) {
                const uint32_t a = lo + 16u * k;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (a + 16 <= len && ((reinterpret_cast<uintptr_t>(src) + a) & 15) == 0) v = ld_nc(reinterpret_cast<const uint4*>(src + a));
                else if (a < len) {
                    uint32_t t[4] = {0, 0, 0, 0};
                    for (uint32_t q = a; q < len && q < a + 16; q++) t[(q - a) >> 2] |= (uint32_t)src[q] << (8 * ((q - a) & 3));
                    v = make_uint4(t[0], t[1], t[2], t[3]);
                }
                reinterpret_cast<uint4*>(sm.win)[k] = v;
            }
            __syncwarp();
        };
        // one input byte: extend the current string if (cur, c) is in the dictionary, else emit cur and add the pair
        auto step = [&](uint32_t c) {
            const uint32_t key = (cur << 8) | c;
            uint32_t h = __umulhi(key * 2654435761u, (uint32_t)kEncSlots);      // multiply-shift range reduction
            uint32_t e = sm.table[h];
            if ((e >> 12) == key) {                      // (an empty slot never matches: no prefix code is 4095)
                cur = e & 0xFFFu;
                return;
            }
            while (e != kEmpty) {                        // linear probing; no deletions, so a present key precedes the first hole
                h = h + 1 == (uint32_t)kEncSlots ? 0u : h + 1;
                e = sm.table[h];
                if ((e >> 12) == key) {
                    cur = e & 0xFFFu;
                    return;
                }
            }
            w.put(cur, nbits, writer);
            sm.table[h] = (key << 12) | (uint32_t)next;  // every lane stores the same word
            next++;
            cur = c;
            if (next == LIMIT) {
                w.put(CLEAR, nbits, writer);
                clear_table();
                nbits = 9;
                next = FIRST;
            } else if (next > (1 << nbits) - 1) {
                nbits++;
            }
        };
        clear_table();
        w.put(CLEAR, nbits, writer);
        if (len) {
            // bytes are taken a staged 32-bit word at a time: word j holds bytes 4j .. 4j+3
            for (uint32_t lo = 0; lo < len; lo += kEncWin) {
                stage(lo);
                const uint32_t words = min((uint32_t)kEncWin, len - lo + 3u) / 4u;
                for (uint32_t j = 0; j < words; j++) {
                    uint32_t word = sm.win[j];
                    const uint32_t b0 = lo + 4u * j;
                    if (b0 >= 1 && b0 + 4 <= len) {      // the common case: four bytes, no edge
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            step(word & 0xFFu);
                            word >>= 8;
                        }
                    } else {
                        for (int k = 0; k < 4; k++) {
                            const uint32_t b = b0 + k;
                            if (b == 0) cur = word & 0xFFu;
                            else if (b < len) step(word & 0xFFu);
                            word >>= 8;
                        }
                    }
                }
            }
            w.put(cur, nbits, writer);
            next++;
            if (next == LIMIT) {
                w.put(CLEAR, nbits, writer);
                nbits = 9;
            } else if (next > (1 << nbits) - 1 && nbits < 12) {
                nbits++;
            }
        }
        w.put(EOI, nbits, writer);
        w.finish(writer);
        if (writer) out_len[si] = w.fail ? 0xFFFFFFFFu : w.o;
        __syncwarp();
    }
}

// (H,W) raster of `pb`-byte pixels -> padded tiles of tw x th pixels, tile-major (row of tiles by row of tiles), zero
// padding on the right / bottom edge: exactly what a tiled, pixel-interleaved TIFF stores per block.
__global__ void __launch_bounds__(256)
tile_split_kernel(const uint8_t* __restrict__ img, int H, int W, int pb, int tw, int th, int across, uint64_t n_out,
                  uint8_t* __restrict__ tiles) {
    const uint64_t row_b = (uint64_t)tw * pb, tile_b = row_b * th;
    for (uint64_t o = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t t = o / tile_b, r = o - t * tile_b;
        const uint32_t ty = (uint32_t)(t / across), tx = (uint32_t)(t - (uint64_t)ty * across);
        const uint32_t y = ty * th + (uint32_t)(r / row_b);
        const uint64_t xb = (uint64_t)tx * row_b + (r % row_b);          // byte column in the image row
        tiles[o] = (y < (uint32_t)H && xb < (uint64_t)W * pb) ? img[(uint64_t)y * W * pb + xb] : (uint8_t)0;
    }
}

// n byte ranges of one device buffer -> back to back (dst_off[i]) in another: the code streams of a batch leave their
// capacity-sized slots for one dense buffer, which then crosses to the host in a single copy.
__global__ void __launch_bounds__(256)
gather_ranges_kernel(const uint8_t* __restrict__ src, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ dst_off,
                     const uint32_t* __restrict__ len, int n, uint8_t* __restrict__ dst) {
    for (int r = blockIdx.y; r < n; r += gridDim.y) {
        const uint8_t* s = src + src_off[r];
        uint8_t* d = dst + dst_off[r];
        const uint32_t ln = len[r];
        // 16-byte loads (the slots are 16-byte aligned), byte stores only at the ragged edges of the destination
        const uint32_t head = min(ln, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(d) & 15u)) & 15u));
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < head; i += gridDim.x * blockDim.x) d[i] = s[i];
        const uint32_t vecs = (ln - head) >> 4;
        for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < vecs; v += gridDim.x * blockDim.x) {
            const uint8_t* p = s + head + 16u * v;                  // source side unaligned in general: four 32-bit pieces
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint8_t* q = p + 4 * k;
                w[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            }
            *reinterpret_cast<uint4*>(d + head + 16u * v) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        for (uint32_t i = head + (vecs << 4) + blockIdx.x * blockDim.x + threadIdx.x; i < ln; i += gridDim.x * blockDim.x) d[i] = s[i];
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_gather_ranges(b2_ctx* ctx, const uint8_t* src, const uint64_t* src_off, const uint64_t* dst_off,
                                const uint32_t* len, int n, uint32_t max_len, uint8_t* dst, b2_stream stream) {
    B2_REQUIRE(ctx && src && src_off && dst_off && len && dst, "b2_gather_ranges: NULL argument");
    if (n <= 0) return 0;
    DeviceGuard g(ctx->device);
    unsigned gx = (max_len + 256 * 64 - 1) / (256 * 64);
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    const unsigned gy = (unsigned)(n < 65535 ? n : 65535);
    gather_ranges_kernel<<<dim3(gx, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, src_off, dst_off, len, n, dst);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b2_lzw_encode(b2_ctx* ctx, const uint8_t* raw, const b2_enc_desc* descs, int n, uint8_t* out,
                             uint32_t* out_len, b2_stream stream) {
    B2_REQUIRE(ctx && raw && descs && out && out_len, "b2_lzw_encode: NULL argument");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b2_lzw_encode: out must be 16-byte aligned (and every dst_off a multiple of 4, every src_len < 2^31)");
    if (n <= 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned ctas = (unsigned)n;
    const unsigned resident = (unsigned)ctx->sm_count * 9;     // 25 KiB of shared memory per CTA
    if (ctas > resident) ctas = resident;
    WsLock ws_lock(ctx);
    if (int e = ws_reserve(ctx, 256, s)) return e;
    unsigned int* counter = static_cast<unsigned int*>(ctx->ws);
    B2_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
    lzw_encode_kernel<<<ctas, 32, 0, s>>>(raw, descs, n, out, out_len, counter);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b2_tile_split(b2_ctx* ctx, const uint8_t* img, int H, int W, int pixel_bytes, int tile_w, int tile_h,
                             uint8_t* tiles, b2_stream stream) {
    B2_REQUIRE(ctx && img && tiles, "b2_tile_split: NULL argument");
    B2_REQUIRE(H >= 1 && W >= 1 && pixel_bytes >= 1 && tile_w >= 1 && tile_h >= 1, "b2_tile_split: bad geometry");
    DeviceGuard g(ctx->device);
    const int across = (W + tile_w - 1) / tile_w, down = (H + tile_h - 1) / tile_h;
    const uint64_t n_out = (uint64_t)across * down * tile_w * tile_h * pixel_bytes;
    uint64_t blocks = (n_out + 255) / 256;
    if (blocks > (uint64_t)ctx->sm_count * 32) blocks = (uint64_t)ctx->sm_count * 32;
    tile_split_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, H, W, pixel_bytes, tile_w, tile_h, across,
                                                                                      n_out, tiles);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}
