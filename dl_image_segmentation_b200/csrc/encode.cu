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
// dictionary built so far): one warp owns one tile, its dictionary is an open-addressing hash table in shared
// memory (8192 slots of { prefix code, byte } -> code), all lanes run the same scalar walk (no divergence) and
// cooperate on what is parallel: clearing the table and staging the input window.  Parallelism comes from the
// thousands of tiles of a batch.
#include "common.cuh"

namespace b2 {

constexpr int kEncWarps = 4;            // streams per CTA
constexpr int kEncSlots = 8192;         // hash slots per stream (4094 entries at most: load < 0.5)
constexpr int kEncWin = 2048;           // staged input bytes per refill
constexpr uint32_t kEmpty = 0xFFFFFFFFu;

// The dictionary of a stream — kEncSlots words of (key << 12) | code, key = (prefix << 8) | byte — lives in the context
// workspace, not in shared memory: the walk is latency-bound (a few hundred cycles per input byte whatever memory the
// table is in), so what counts is how many streams are resident, and 32 KiB of shared memory per stream would allow
// six per SM where the L2-resident tables allow sixty-four.
struct EncSmem {
    uint8_t win[kEncWin];
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

__global__ void __launch_bounds__(kEncWarps * 32)
lzw_encode_kernel(const uint8_t* __restrict__ raw, const b2_enc_desc* __restrict__ descs, int n, uint8_t* __restrict__ out,
                  uint32_t* __restrict__ out_len, unsigned int* next_stream, uint32_t* tables) {
    __shared__ __align__(16) EncSmem sm_all[kEncWarps];
    EncSmem* sm = &sm_all[threadIdx.x >> 5];
    uint32_t* table = tables + (size_t)(blockIdx.x * kEncWarps + (threadIdx.x >> 5)) * kEncSlots;
    const int lane = threadIdx.x & 31;
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
        for (int i = lane; i < kEncSlots; i += 32) __stcg(table + i, kEmpty);
        uint32_t win_lo = 0x80000000u;                   // window = [win_lo, win_lo + kEncWin); the first access misses
        auto byte_at = [&](uint32_t p) -> uint32_t {
            if (p - win_lo >= (uint32_t)kEncWin) {       // warp-uniform: refill with coalesced 16-byte loads
                __syncwarp();
                win_lo = p & ~15u;
                for (int k = lane; k < kEncWin / 16; k += 32) {
                    const uint32_t a = win_lo + 16u * k;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (a + 16 <= len && ((reinterpret_cast<uintptr_t>(src) + a) & 15) == 0) v = ld_nc(reinterpret_cast<const uint4*>(src + a));
                    else if (a < len) {
                        uint32_t t[4] = {0, 0, 0, 0};
                        for (uint32_t q = a; q < len && q < a + 16; q++) t[(q - a) >> 2] |= (uint32_t)src[q] << (8 * ((q - a) & 3));
                        v = make_uint4(t[0], t[1], t[2], t[3]);
                    }
                    reinterpret_cast<uint4*>(sm->win)[k] = v;
                }
                __syncwarp();
            }
            return sm->win[p - win_lo];
        };
        __syncwarp();
        w.put(CLEAR, nbits, writer);
        if (len) {
            uint32_t cur = byte_at(0);
            for (uint32_t i = 1; i < len; i++) {
                const uint32_t c = byte_at(i);
                const uint32_t key = (cur << 8) | c;
                uint32_t h = (key * 2654435761u) >> 19;  // 13 bits
                uint32_t found = kEmpty;
                for (;;) {
                    const uint32_t e = __ldcg(table + h);
                    if (e == kEmpty) break;
                    if ((e >> 12) == key) { found = e & 0xFFFu; break; }
                    h = (h + 1) & (kEncSlots - 1);
                }
                if (found != kEmpty) {
                    cur = found;
                    continue;
                }
                w.put(cur, nbits, writer);
                if (writer) __stcg(table + h, (key << 12) | (uint32_t)next);
                __syncwarp();
                next++;
                cur = c;
                if (next == LIMIT) {
                    w.put(CLEAR, nbits, writer);
                    __syncwarp();
                    for (int k = lane; k < kEncSlots; k += 32) __stcg(table + k, kEmpty);
                    __syncwarp();
                    nbits = 9;
                    next = FIRST;
                } else if (next > (1 << nbits) - 1) {
                    nbits++;
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

}  // namespace b2

using namespace b2;

extern "C" int b2_lzw_encode(b2_ctx* ctx, const uint8_t* raw, const b2_enc_desc* descs, int n, uint8_t* out,
                             uint32_t* out_len, b2_stream stream) {
    B2_REQUIRE(ctx && raw && descs && out && out_len, "b2_lzw_encode: NULL argument");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b2_lzw_encode: out must be 16-byte aligned (and every dst_off a multiple of 4, every src_len < 2^31)");
    if (n <= 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    unsigned ctas = (unsigned)((n + kEncWarps - 1) / kEncWarps);
    const unsigned resident = (unsigned)ctx->sm_count * (64 / kEncWarps);
    if (ctas > resident) ctas = resident;
    const size_t table_bytes = (size_t)ctas * kEncWarps * kEncSlots * sizeof(uint32_t);
    if (int e = ws_reserve(ctx, table_bytes + 256, s)) return e;
    unsigned int* counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(ctx->ws) + table_bytes);
    B2_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
    lzw_encode_kernel<<<ctas, kEncWarps * 32, 0, s>>>(raw, descs, n, out, out_len, counter, static_cast<uint32_t*>(ctx->ws));
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
