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
#include <algorithm>
#include <vector>

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

// ---------------------------------------------------------------- restart-interval encoder
// TIFF-LZW allows a Clear code anywhere.  With a Clear every R input bytes (R <= 1024) the pieces of a tile between two
// Clears are independent LZW streams: a tile of 512 KiB is 512 of them instead of one serial walk, the dictionary of a
// piece has at most R entries, so its hash table is 8 KiB instead of 24 (28 of them per SM, each owned by ONE thread —
// the serial walk has nothing for the other lanes of a warp to do), and codes stay 9-11 bits wide.  Two kernels:
//   lzw_segment_kernel  the serial part, as short as it gets: probe, and on a miss store the code as a 16-bit number and
//                       insert.  No bit packing here: the width of the i-th code of a piece depends on i alone.
//   lzw_pack_kernel     per tile: bit lengths of the pieces from their code counts (closed form), exclusive scan, then every
//                       thread packs eight consecutive 32-bit words of the final stream from the 16-bit codes.
// Two lanes per warp run a piece each (14 warps per CTA).  The walk is a chain of dependent instructions, ~35 per input byte
// on noisy imagery (96 % of the steps are misses): with 7 lanes in each of 4 warps the lanes execute each other's probe
// loops and miss paths and one warp per scheduler hides nothing (latency-bound, 410 cycles per byte-step, 19 GB/s); with
// one lane in each of 28 warps the chains overlap but every instruction serves one byte (issue-bound at 79 %, 21 GB/s);
// 2 x 14 measured best (26 GB/s; 4 x 7: 24.7).
constexpr int kSegSlots = 2048;         // hash slots per piece: load <= 0.5
constexpr int kSegLanes = 2;            // active lanes per warp ...
constexpr int kSegWarps = 14;           // ... times warps = 28 private tables = 224 KiB of shared memory
constexpr int kSegMaxRestart = 1024;

struct SegArgs {
    const uint8_t* raw;
    const b2_enc_desc* descs;           // device copy, all n tiles
    const uint32_t* seg_start;          // [n + 1] first piece of every tile (global numbering)
    uint32_t tile0, tile1;              // this launch covers the pieces of tiles [tile0, tile1)
    uint32_t seg0, seg1;                // = seg_start[tile0], seg_start[tile1]
    uint32_t restart, slot_codes;
    uint16_t* scratch;                  // (seg1 - seg0) slots of slot_codes codes
    uint32_t* seg_count;                // codes per piece
    uint32_t* seg_pos;                  // exclusive scan of the pieces' bit lengths inside every tile
    unsigned int* counter;
    uint8_t* out;
    uint32_t* out_len;
};

__global__ void __launch_bounds__(kSegWarps * 32, 1)
lzw_segment_kernel(const SegArgs a) {
    extern __shared__ __align__(16) uint32_t seg_tables[];
    if ((threadIdx.x & 31) >= kSegLanes) return;
    uint32_t* tab = seg_tables + (size_t)((threadIdx.x >> 5) * kSegLanes + (threadIdx.x & 31)) * kSegSlots;
    for (;;) {
        const uint32_t seg = a.seg0 + atomicAdd(a.counter, 1u);
        if (seg >= a.seg1) return;
        // the tile owning this piece: last t in [tile0, tile1) with seg_start[t] <= seg
        uint32_t lo_t = a.tile0, hi_t = a.tile1;
        while (hi_t - lo_t > 1) {
            const uint32_t mid = (lo_t + hi_t) >> 1;
            if (__ldg(&a.seg_start[mid]) <= seg) lo_t = mid; else hi_t = mid;
        }
        const uint32_t k = seg - __ldg(&a.seg_start[lo_t]);
        const b2_enc_desc d = a.descs[lo_t];
        const uint32_t lo = k * a.restart;
        const uint32_t cnt = d.src_len > lo ? min(a.restart, d.src_len - lo) : 0u;
        const uint8_t* p = a.raw + d.src_off + lo;
        uint16_t* codes = a.scratch + (size_t)(seg - a.seg0) * a.slot_codes;
        for (int q = 0; q < kSegSlots / 4; q++) reinterpret_cast<uint4*>(tab)[q] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
        uint32_t next = 258, cur = 0;                                   // next dictionary code = 258 + codes emitted
        uint16_t* cp = codes;
        const bool aligned = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
        auto load16 = [&](uint32_t off) {
            if (aligned && off + 16 <= cnt) return ld_nc(reinterpret_cast<const uint4*>(p + off));
            uint32_t t[4] = {0, 0, 0, 0};
            for (uint32_t q = off; q < cnt && q < off + 16; q++) t[(q - off) >> 2] |= (uint32_t)p[q] << (8 * ((q - off) & 3));
            return make_uint4(t[0], t[1], t[2], t[3]);
        };
        auto step = [&](uint32_t c) {
            const uint32_t key = (cur << 8) | c;
            uint32_t h = (key * 2654435761u) >> 21;                     // top 11 bits: 0 .. kSegSlots - 1
            uint32_t e = tab[h];
            while (e != kEmpty && (e >> 11) != key) {
                h = (h + 1) & (kSegSlots - 1);
                e = tab[h];
            }
            if (e != kEmpty) {
                cur = e & 0x7FFu;
            } else {
                *cp++ = (uint16_t)cur;
                tab[h] = (key << 11) | next;
                next++;
                cur = c;
            }
        };
        if (cnt) {
            uint4 nxt = load16(0);
            for (uint32_t off = 0; off < cnt; off += 16) {
                const uint4 v = nxt;
                if (off + 16 < cnt) nxt = load16(off + 16);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                if (off != 0 && off + 16 <= cnt) {
#pragma unroll
                    for (int q = 0; q < 16; q++) step((w[q >> 2] >> (8 * (q & 3))) & 0xFFu);
                } else {
                    const uint32_t m = min(16u, cnt - off);
                    for (uint32_t q = 0; q < m; q++) {
                        const uint32_t c = (w[q >> 2] >> (8 * (q & 3))) & 0xFFu;
                        if (off + q == 0) cur = c;
                        else step(c);
                    }
                }
            }
            *cp++ = (uint16_t)cur;
        }
        a.seg_count[seg - a.seg0] = (uint32_t)(cp - codes);
    }
}

// Widths inside a piece: the dictionary holds 258 + i codes when code i goes out, and the width grows one code early
// ("early change"): codes 0..253 have 9 bits, 254..765 ten, from 766 eleven (a piece has at most 1025 codes).
__device__ __forceinline__ uint32_t seg_code_bits(uint32_t i) { return i < 254 ? 9u : (i < 766 ? 10u : 11u); }
__device__ __forceinline__ uint32_t seg_bits_before(uint32_t i) {      // bits of codes 0 .. i-1
    return i <= 254 ? 9u * i : (i <= 766 ? 2286u + 10u * (i - 254) : 7406u + 11u * (i - 766));
}
__device__ __forceinline__ uint32_t seg_code_at(uint32_t x) {          // index of the code holding bit x
    return x < 2286u ? x / 9u : (x < 7406u ? 254u + (x - 2286u) / 10u : 766u + (x - 7406u) / 11u);
}
// a piece in the final stream: [Clear, 9 bits, first piece of a tile only] codes [Clear | EOI in the width code n would have]
__device__ __forceinline__ uint32_t seg_total_bits(uint32_t n, bool first) {
    return (first ? 9u : 0u) + seg_bits_before(n) + seg_code_bits(n);
}

constexpr int kCatThreads = 256;
constexpr int kCatWords = 8;            // consecutive output words per thread (one 32-byte sector)

__global__ void __launch_bounds__(kCatThreads)
lzw_pack_kernel(const SegArgs a) {
    __shared__ uint32_t warp_sum[kCatThreads / 32];
    __shared__ uint32_t carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t t = a.tile0 + blockIdx.x; t < a.tile1; t += gridDim.x) {
        const uint32_t s0 = a.seg_start[t] - a.seg0, nseg = a.seg_start[t + 1] - a.seg_start[t];
        const uint32_t* count = a.seg_count + s0;
        uint32_t* pos = a.seg_pos + s0;
        __syncthreads();
        if (tid == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < nseg; base += kCatThreads) {      // exclusive scan, 256 pieces at a time
            const uint32_t i = base + tid;
            const uint32_t v = i < nseg ? seg_total_bits(count[i], i == 0) : 0u;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) warp_sum[warp] = x;
            __syncthreads();
            uint32_t before = carry;
            for (int q = 0; q < warp; q++) before += warp_sum[q];
            if (i < nseg) pos[i] = before + x - v;
            __syncthreads();
            if (tid == kCatThreads - 1) carry = before + x;
            __syncthreads();
        }
        const uint32_t total = carry;
        const b2_enc_desc d = a.descs[t];
        const uint32_t nbytes = (total + 7) >> 3;
        if (nbytes > d.dst_cap) {
            if (tid == 0) a.out_len[t] = 0xFFFFFFFFu;
            continue;
        }
        if (tid == 0) a.out_len[t] = nbytes;
        uint8_t* dst = a.out + d.dst_off;
        const uint32_t nwords = (total + 31) >> 5;
        for (uint32_t j0 = (uint32_t)tid * kCatWords; j0 < nwords; j0 += kCatThreads * kCatWords) {
            // the piece holding bit 32 * j0: last k with pos[k] <= bit
            const uint32_t b0 = j0 << 5;
            uint32_t lo = 0, hi = nseg;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (pos[mid] <= b0) lo = mid; else hi = mid;
            }
            uint32_t k = lo, n = count[k];
            const uint16_t* codes = a.scratch + (size_t)(s0 + k) * a.slot_codes;
            // virtual code index inside the piece: -1 = the tile's leading Clear, 0 .. n-1 the stored codes, n the Clear / EOI
            int i;
            uint32_t skip;                                              // bits of that code already in earlier words
            {
                uint32_t off = b0 - pos[k];
                if (k == 0) {
                    if (off < 9) { i = -1; skip = off; }
                    else { off -= 9; i = (int)min(seg_code_at(off), n); skip = off - seg_bits_before((uint32_t)i); }
                } else {
                    i = (int)min(seg_code_at(off), n);
                    skip = off - seg_bits_before((uint32_t)i);
                }
            }
            unsigned long long acc = 0;
            uint32_t nacc = 0, made = 0;
            uint32_t vals[kCatWords];
#pragma unroll
            for (int q = 0; q < kCatWords; q++) vals[q] = 0;
            while (made < kCatWords && k < nseg) {
                uint32_t val, width;
                if (i < 0) { val = 256; width = 9; }
                else if ((uint32_t)i < n) { val = codes[i]; width = seg_code_bits((uint32_t)i); }
                else { val = k + 1 == nseg ? 257u : 256u; width = seg_code_bits(n); }
                if (skip) {
                    width -= skip;
                    val &= (1u << width) - 1u;
                    skip = 0;
                }
                acc = (acc << width) | val;
                nacc += width;
                if (nacc >= 32) {
                    const uint32_t wv = (uint32_t)(acc >> (nacc - 32));
#pragma unroll
                    for (int q = 0; q < kCatWords; q++)
                        if (q == (int)made) vals[q] = wv;
                    made++;
                    nacc -= 32;
                }
                if (i >= 0 && (uint32_t)i >= n) {                       // past this piece's last code: the next piece
                    k++;
                    if (k < nseg) {
                        n = count[k];
                        codes += a.slot_codes;
                        i = 0;
                    }
                } else {
                    i++;
                }
            }
            if (made < kCatWords && nacc) {                             // the end of the stream: zero padding
                const uint32_t wv = (uint32_t)(acc << (32 - nacc));
#pragma unroll
                for (int q = 0; q < kCatWords; q++)
                    if (q == (int)made) vals[q] = wv;
                made++;
            }
#pragma unroll
            for (int q = 0; q < kCatWords; q++) {
                const uint32_t j = j0 + q;
                if (q >= (int)made || j >= nwords) break;
                const uint32_t be = __byte_perm(vals[q], 0, 0x0123);    // big-endian bit stream
                if (4 * j + 4 <= d.dst_cap) *reinterpret_cast<uint32_t*>(dst + 4 * (size_t)j) = be;
                else
                    for (uint32_t z = 4 * j; z < nbytes; z++) dst[z] = (uint8_t)(be >> (8 * (z - 4 * j)));
            }
        }
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

extern "C" int b2_lzw_encode_restart(b2_ctx* ctx, const uint8_t* raw, const b2_enc_desc* descs_host, int n, uint32_t restart_bytes,
                                     uint8_t* out, uint32_t* out_len, b2_stream stream) {
    B2_REQUIRE(ctx && raw && descs_host && out && out_len, "b2_lzw_encode_restart: NULL argument");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b2_lzw_encode_restart: out must be 16-byte aligned (and every dst_off a multiple of 4)");
    B2_REQUIRE(restart_bytes >= 16 && restart_bytes <= (uint32_t)kSegMaxRestart && restart_bytes % 16 == 0,
               "b2_lzw_encode_restart: restart_bytes must be a multiple of 16 in [16, 1024]");
    if (n <= 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    std::vector<uint32_t> seg_start((size_t)n + 1);
    uint64_t segs = 0;
    for (int i = 0; i < n; i++) {
        B2_REQUIRE(descs_host[i].dst_off % 4 == 0, "b2_lzw_encode_restart: dst_off must be a multiple of 4");
        B2_REQUIRE(descs_host[i].src_len < (1u << 28), "b2_lzw_encode_restart: src_len must be below 2^28");
        seg_start[i] = (uint32_t)segs;
        const uint64_t k = ((uint64_t)descs_host[i].src_len + restart_bytes - 1) / restart_bytes;
        segs += k ? k : 1;
        B2_REQUIRE(segs < (1ull << 31), "b2_lzw_encode_restart: too many pieces for one call");
    }
    seg_start[n] = (uint32_t)segs;
    // a piece has at most restart codes (one per input byte), stored as 16-bit numbers; slots stay 16-byte aligned
    const uint32_t slot_codes = restart_bytes + 8;
    // tiles are encoded in groups whose scratch stays under 256 MiB
    const uint64_t budget_segs = std::max<uint64_t>(1, (256ull << 20) / (slot_codes * 2ull));
    uint64_t group_max = 0;
    for (int t0 = 0; t0 < n;) {
        int t1 = t0 + 1;
        while (t1 < n && seg_start[t1 + 1] - seg_start[t0] <= budget_segs) t1++;
        group_max = std::max<uint64_t>(group_max, seg_start[t1] - seg_start[t0]);
        t0 = t1;
    }
    const size_t off_descs = 256, off_start = off_descs + (((size_t)n * sizeof(b2_enc_desc) + 255) & ~(size_t)255);
    const size_t off_bits = off_start + (((size_t)(n + 1) * 4 + 255) & ~(size_t)255);
    const size_t off_pos = off_bits + ((group_max * 4 + 255) & ~(size_t)255);
    const size_t off_scratch = off_pos + ((group_max * 4 + 255) & ~(size_t)255);
    WsLock ws_lock(ctx);
    if (int e = ws_reserve(ctx, off_scratch + group_max * slot_codes * 2ull + 256, s)) return e;
    uint8_t* ws = static_cast<uint8_t*>(ctx->ws);
    B2_CUDA(cudaMemcpyAsync(ws + off_descs, descs_host, (size_t)n * sizeof(b2_enc_desc), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(ws + off_start, seg_start.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaStreamSynchronize(s));                                  // the host arrays are read by the copies above
    const int smem = kSegWarps * kSegLanes * kSegSlots * 4;
    // (per device: a process may hold one context per GPU, so no "done once" flag)
    B2_CUDA(cudaFuncSetAttribute(lzw_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SegArgs a{raw, reinterpret_cast<const b2_enc_desc*>(ws + off_descs), reinterpret_cast<const uint32_t*>(ws + off_start), 0, 0, 0, 0,
              restart_bytes, slot_codes, reinterpret_cast<uint16_t*>(ws + off_scratch), reinterpret_cast<uint32_t*>(ws + off_bits),
              reinterpret_cast<uint32_t*>(ws + off_pos), reinterpret_cast<unsigned int*>(ws), out, out_len};
    for (int t0 = 0; t0 < n;) {
        int t1 = t0 + 1;
        while (t1 < n && seg_start[t1 + 1] - seg_start[t0] <= budget_segs) t1++;
        a.tile0 = (uint32_t)t0;
        a.tile1 = (uint32_t)t1;
        a.seg0 = seg_start[t0];
        a.seg1 = seg_start[t1];
        B2_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(unsigned int), s));
        const uint32_t pieces = a.seg1 - a.seg0;
        unsigned ctas = (pieces + kSegWarps * kSegLanes - 1) / (kSegWarps * kSegLanes);
        if (ctas > (unsigned)ctx->sm_count) ctas = (unsigned)ctx->sm_count;
        lzw_segment_kernel<<<ctas, kSegWarps * 32, smem, s>>>(a);
        unsigned cat = (unsigned)(t1 - t0);
        if (cat > (unsigned)ctx->sm_count * 8) cat = (unsigned)ctx->sm_count * 8;
        lzw_pack_kernel<<<cat, kCatThreads, 0, s>>>(a);
        ctx->launches += 2;
        B2_CUDA(cudaGetLastError());
        t0 = t1;
    }
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
