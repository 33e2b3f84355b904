// codec.cu — K1: chip decode on the GPU.  TIFF-LZW, DEFLATE (zlib-wrapped: TIFF compression 8/32946 and PNG
// IDAT), PNG un-filter, TIFF predictor 2 / byte order / planar / tile assembly; plus the host-side header
// parsers (TIFF IFD, PNG chunks) that plan the device work.
//
// Replaces (reference call sites): rasterio MemoryFile(...).open().read() -> GDAL -> libtiff / libpng
//     _img_to_tf_mp.py:45-48, _tfrecord_image_translation.py:320-326,369-381
//   tf.image.decode_png / tf.io.decode_image -> libpng + zlib
//     _img_to_tf_threaded.py:59, _tfrecord_image_translation.py:283,289
//   reshape_as_image (B,H,W)->(H,W,B)  _img_to_tf_mp.py:69  (output is written HWC directly)
// Chip format: _descartes_img_chips.py:781-797 (GTiff, COMPRESS=LZW, TILED=TRUE).
//
// Entropy decode is serial inside a stream, so parallelism is one WARP per compressed stream (tile, strip or PNG
// zlib stream) with thousands of streams in flight, and the warp's 32 lanes are used inside the stream:
//   LZW    : 32 codes per round.  Inside a Clear-delimited segment the code widths depend only on the code index,
//            so all 32 bit positions are known up front; string lengths resolve through the "entry k = output of
//            code k plus one byte" identity (LZW is LZ77 with implicit back references), a warp scan gives the
//            output offsets, and the bytes are produced in parallel by chasing each byte to its literal.
//   DEFLATE: every lane decodes the same Huffman symbols (uniform control flow, no broadcasts), literals are
//            gathered 32 at a time into one coalesced store, and LZ77 matches are copied by the whole warp.
//   PNG    : anti-diagonal wavefront, one lane per scanline of a 32-row band (left / up / up-left dependencies).
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <cstring>
#include <vector>

#include "common.cuh"

namespace b2 {

enum { CODEC_RAW = 1, CODEC_LZW = 5, CODEC_ZLIB = 8 };

__device__ __forceinline__ void set_status(int32_t* status, int image, int code) {
    if (status) atomicMax(status + image, code);
}

// ================================================================================================ LZW
// bits consumed by the first k codes of a segment (k = number of codes after the Clear)
// (branch-free: every code costs 9 bits, plus one more for each width step it lies beyond)
__device__ __forceinline__ uint32_t lzw_cum_bits(uint32_t k) {
    const int s = (int)k;
    return 9u * k + (uint32_t)(max(s - 254, 0) + max(s - 766, 0) + max(s - 1790, 0));
}
__device__ __forceinline__ uint32_t lzw_width(uint32_t k) { return 9u + (k >= 254) + (k >= 766) + (k >= 1790); }

// One CTA per compressed stream (tile or strip), a whole Clear-delimited SEGMENT per pass (<= 3838 codes):
//   A  the segment's input words are staged in shared memory with coalesced loads; thread t then extracts codes
//      15 t .. 15 t + 14 from a sliding 64-bit window (a code's bit position depends only on its index); the first control
//      code (Clear / EOI / end of input) bounds the segment;
//   B  code k >= 258 names table entry e = code - 258 = "string of code e plus the first byte of string e + 1", so the
//      strings form a forest over the code positions (parent(k) = e < k).  Pointer jumping over that forest gives every
//      string's depth (= length - 1) and root (= its first byte) in O(log depth) rounds, each thread keeping its 15 nodes
//      in registers;
//   C  a block scan of the lengths gives the output offsets (registers again: a thread emits its own strings);
//   D  one thread per string walks its chain and writes the bytes back to front — the last byte of string k is the
//      first byte of string e + 1, then the same for e's own entry, ... down to the root literal: the classic table walk,
//      except that no table is ever built and every string of the segment is produced at the same time.  The bytes go to
//      a shared-memory window laid out with the alignment of their destination and leave in 16-byte stores.
// About 1.5 warp instructions per output byte on noisy 16-bit chips (the 32-codes-per-warp-round kernel this replaces
// needed about eight).
constexpr int kLzwThreads = 256;
constexpr int kLzwPer = 15;                          // codes per thread and segment
constexpr int kLzwMaxCodes = kLzwThreads * kLzwPer;  // 3840 >= 3838 = entries 258..4095 plus the code that sees the table full
constexpr int kLzwInWords = 1408;                    // 43 270 bits of a full segment + alignment, rounded to 5.5 words per thread
constexpr int kLzwWindow = 12288;                    // output bytes per flush

struct LzwSmem {
    uint32_t tcode[kLzwMaxCodes + 8];  // [15:0] code at position k, [23:16] first byte of its string (after phase B)
    union {
        uint32_t tnode[kLzwMaxCodes + 8];  // [15:0] pointer-jumping ancestor, [31:16] hops to it (phases B, C)
        uint4 out16[kLzwWindow / 16 + 2];  // output window (phase D)
    } u;
    uint32_t in[2][kLzwInWords];       // the segment's input words as they lie in memory; the next segment's arrive meanwhile
    uint32_t warp_sum[kLzwThreads / 32];
    int m[2];                          // position of the first control code: [0] of a full pass, [1] of a short one (a short
                                       // pass that gives up is followed at once, without a barrier, by a full pass)
    int stream;                        // next stream drawn from the counter
    // the stream being decoded (written once per stream; kept out of the registers of the segment pass)
    const uint32_t* wsrc;              // its words (aligned down); stream bit 0 = bit 8 * mis of word 0
    uint8_t* dst;
    uint32_t n_words, mis, src_bits, dst_len;
};
static_assert(sizeof(LzwSmem) <= 48 * 1024, "lzw_kernel uses static shared memory");

// stage words [w0 + j0, w0 + j1) of the stream into sdst[j0 .. j1) (zeros beyond its last word) with asynchronous 4-byte copies
__device__ __forceinline__ void lzw_stage(uint32_t* sdst, const uint32_t* __restrict__ wsrc, uint32_t w0, uint32_t n_words, int tid,
                                          int j0, int j1) {
    for (int j = j0 + tid; j < j1; j += kLzwThreads) {
        const uint32_t w = w0 + (uint32_t)j;
        const uint32_t ok = w < n_words ? 4u : 0u;
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst + j);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(wsrc + (ok ? w : 0u)), "r"(ok) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// codes kPer t .. kPer t + kPer - 1 of the segment from a sliding two-word window; kCheck: the input may end inside this span
template <int kPer, bool kCheck>
__device__ __forceinline__ uint32_t lzw_extract(const uint32_t* __restrict__ in, uint32_t bit0, uint32_t wd0, int step_at,
                                                uint32_t lim, uint32_t (&code)[kPer], uint32_t* tcode) {
    uint32_t widx = bit0 >> 5, sh = bit0 & 31u;
    uint32_t hi = __byte_perm(in[widx], 0u, 0x0123), lo = __byte_perm(in[widx + 1], 0u, 0x0123);
    widx += 2;
    uint32_t ctl = 0, used = 0;
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        const uint32_t wd = wd0 + (i >= step_at ? 1u : 0u);
        uint32_t c = __funnelshift_l(lo, hi, sh) >> (32u - wd);
        sh += wd;
        if (sh >= 32u) { sh -= 32u; hi = lo; lo = __byte_perm(in[widx++], 0u, 0x0123); }
        if (kCheck) {
            used += wd;
            if (used > lim) c = 257u;                                  // running out of input behaves like EOI
        }
        code[i] = c;
        tcode[i] = c;
        if ((c >> 1) == 128u) ctl |= 1u << i;                          // Clear (256) or EOI (257)
    }
    return ctl;
}

// Segments come in two sizes.  A full dictionary ends a segment after 3838 codes (15 per thread); a writer that restarts
// the dictionary every 1024 input bytes (b2_lzw_encode_restart, the GeoTIFF writer of this package) makes segments of at
// most 1026 codes, and a pass costs about the same whatever the segment holds — so there is a second instantiation of the
// pass with 5 codes per thread (1280 positions), tried first and again after every segment that would have fitted it.
constexpr int kLzwPerShort = 5;
constexpr int kLzwShortWords = 512;                  // 1280 codes x 12 bits + alignment

struct LzwStream {
    uint32_t bitpos, out;                            // consumed input bits, produced output bytes
    int err, buf;
    int staged;                                      // words of the current segment staged in sm.in[buf]
    bool next_short;                                 // the pass to try on the next segment
};

// One segment.  Returns 0: done, the next segment follows; 1: the stream ends here (EOI, end of input, error: st.err);
// 2 (short pass only): no control code among the first 1280 codes — nothing was consumed, the long pass must take it.
template <int kPer>
__device__ __forceinline__ int lzw_pass(LzwSmem& sm, LzwStream& st, const int tid, const int lane, const int warp) {
    constexpr int kMax = kPer * kLzwThreads;
    const int k0 = tid * kPer;
    // code width of this thread's positions: 9 bits, one more from positions 254, 766 and 1790 on; a thread's positions
    // cross at most one of the three steps
    const uint32_t wd0 = lzw_width((uint32_t)k0);
    const int next_step = k0 < 254 ? 254 : (k0 < 766 ? 766 : (k0 < 1790 ? 1790 : (1 << 30)));
    const int step_at = next_step - k0;
    const uint32_t rel0 = lzw_cum_bits((uint32_t)k0);                  // bits from the segment start to this thread's first code
    uint8_t* const sout = reinterpret_cast<uint8_t*>(sm.u.out16);
    const uint32_t mis = sm.mis, bitpos = st.bitpos, src_bits = sm.src_bits;
    const int buf = st.buf;
    // ---- A: all codes of the segment (its input was staged while the previous segment was being written out)
    const uint32_t r0 = (8u * mis + bitpos) & 31u;
    int* const first_ctl = &sm.m[kPer == kLzwPerShort ? 1 : 0];
    if (tid == 0) *first_ctl = kMax;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    uint32_t code[kPer];
    {
        const uint32_t remaining = src_bits - bitpos;              // bits of the stream from the segment start on
        const uint32_t lim = remaining > rel0 ? remaining - rel0 : 0u;
        const uint32_t ctl = (lim >= 12u * kPer)
            ? lzw_extract<kPer, false>(sm.in[buf], r0 + rel0, wd0, step_at, lim, code, sm.tcode + k0)
            : lzw_extract<kPer, true>(sm.in[buf], r0 + rel0, wd0, step_at, lim, code, sm.tcode + k0);
        if (ctl) atomicMin(first_ctl, k0 + __ffs(ctl) - 1);
    }
    __syncthreads();
    const int m = *first_ctl;
    if (kPer == kLzwPerShort && m == kMax) return 2;
    {   // the next segment starts right after this one's Clear: fetch its input now
        const uint32_t nb = bitpos + lzw_cum_bits((uint32_t)m) + lzw_width((uint32_t)m);
        st.next_short = m < kLzwPerShort * kLzwThreads - 2;
        st.staged = st.next_short ? kLzwShortWords : kLzwInWords;
        if (m < kMax && nb < src_bits) lzw_stage(sm.in[buf ^ 1], sm.wsrc, (8u * mis + nb) >> 5, sm.n_words, tid, 0, st.staged);
    }
    // ---- B: forest over the code positions -> depth and root by pointer jumping (own nodes in registers)
    uint32_t node[kPer];
    bool bad = false;
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        const uint32_t k = (uint32_t)(k0 + i), c = code[i];
        uint32_t nd = k;                                           // literal / beyond the segment: a root
        if ((int)k < m && c >= 256u) {
            const uint32_t e = c - 258u;
            if (e >= k) bad = true;                                // entry not defined yet (k == 0: the first code must be a literal)
            else nd = e | (1u << 16);
        }
        node[i] = nd;
        sm.u.tnode[k] = nd;
    }
    if (__syncthreads_or(bad)) { st.err = 2; return 1; }
    for (;;) {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const uint32_t j = node[i] & 0xFFFFu, par = sm.u.tnode[j];
            if (par != j) {                                        // the ancestor is not a root yet: hop over it
                node[i] = ((node[i] & 0xFFFF0000u) + (par & 0xFFFF0000u)) | (par & 0xFFFFu);
                changed = true;
            }
        }
        if (!__syncthreads_or(changed)) break;                     // every ancestor is a root: depths are final
#pragma unroll
        for (int i = 0; i < kPer; i++) sm.u.tnode[k0 + i] = node[i];
        __syncthreads();
    }
    // ---- C: first bytes, lengths, offsets
    uint32_t run = 0;
#pragma unroll
    for (int i = 0; i < kPer; i++) {
        if (k0 + i < m) {
            if (code[i] >= 256u) {
                const uint32_t fb = sm.tcode[node[i] & 0xFFFFu] & 0xFFu;   // the root is a literal; its low byte never changes
                sm.tcode[k0 + i] = code[i] | (fb << 16);
            } else {
                sm.tcode[k0 + i] = code[i] * 0x10001u;
            }
            run += (node[i] >> 16) + 1u;
        }
    }
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm.warp_sum[warp] = incl;
    __syncthreads();                                               // also: tnode is dead, its space becomes the output window
    uint32_t my_off = incl - run, total = 0;
#pragma unroll
    for (int w = 0; w < kLzwThreads / 32; w++) {
        const uint32_t v = sm.warp_sum[w];
        if (w < warp) my_off += v;
        total += v;
    }
    // ---- D: every string, back to front, through an output window that shares the destination's 16-byte alignment.
    // libtiff truncates the last string of a tile: bytes past dst_len are dropped
    const uint32_t room = sm.dst_len - st.out;
    const uint32_t emit = total < room ? total : room;
    uint8_t* const gseg = sm.dst + st.out;
    if (emit == total && total <= (uint32_t)kLzwWindow) {
        // the whole segment fits one window and the tile (the common case): no range checks on the way
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(gseg) & 15u);
        uint8_t* ptr = sout + shift + my_off;
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            if (k0 + i < m) {
                uint32_t c = code[i];
                if (c < 256u) {
                    *ptr++ = (uint8_t)c;
                } else {
                    ptr += (node[i] >> 16) + 1u;
                    uint8_t* q = ptr;
                    do {
                        *--q = (uint8_t)(sm.tcode[c - 257u] >> 16);
                        c = sm.tcode[c - 258u] & 0xFFFFu;
                    } while (c >= 256u);
                    *--q = (uint8_t)c;
                }
            }
        }
        __syncthreads();
        uint8_t* const g16 = gseg - shift;                         // 16-byte aligned
        const uint32_t lo = shift, hi = shift + total;             // valid bytes of the window buffer
        for (uint32_t c16 = (uint32_t)tid; c16 * 16u < hi; c16 += kLzwThreads) {
            const uint32_t b0 = c16 * 16u;
            if (b0 >= lo && b0 + 16u <= hi) {
                *reinterpret_cast<uint4*>(g16 + b0) = sm.u.out16[c16];
            } else {
                for (uint32_t q = (b0 > lo ? b0 : lo); q < b0 + 16u && q < hi; q++) g16[q] = sout[q];
            }
        }
        __syncthreads();
    } else
    for (uint32_t wv = 0; wv < emit; wv += kLzwWindow) {
        const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(gseg + wv) & 15u);
        const uint32_t wend = (wv + kLzwWindow < emit) ? wv + kLzwWindow : emit;
        uint8_t* const sbase = sout + shift - wv;                  // window byte of segment byte p: sbase[p]
        uint32_t o = my_off;
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            if (k0 + i < m) {
                const uint32_t c0 = code[i];
                if (c0 < 256u) {                                   // a literal: one byte
                    if (o >= wv && o < wend) sbase[o] = (uint8_t)c0;
                    o += 1u;
                } else {
                    const uint32_t len = (node[i] >> 16) + 1u;
                    if (o < wend && o + len > wv) {
                        uint32_t p = o + len - 1u, c = c0;
                        for (;;) {
                            uint32_t byte = c;
                            if (c >= 256u) {
                                byte = (sm.tcode[c - 257u] >> 16) & 0xFFu;
                                c = sm.tcode[c - 258u] & 0xFFFFu;
                            } else {
                                c = 0xFFFFFFFFu;
                            }
                            if (p < wend) sbase[p] = (uint8_t)byte;
                            if (c == 0xFFFFFFFFu || p <= wv) break;
                            p--;
                        }
                    }
                    o += len;
                }
            }
        }
        __syncthreads();
        {
            uint8_t* const g16 = gseg + wv - shift;                // 16-byte aligned
            const uint32_t lo = shift, hi = shift + (wend - wv);   // valid bytes of the window buffer
            for (uint32_t c16 = (uint32_t)tid; c16 * 16u < hi; c16 += kLzwThreads) {
                const uint32_t b0 = c16 * 16u;
                if (b0 >= lo && b0 + 16u <= hi) {
                    *reinterpret_cast<uint4*>(g16 + b0) = sm.u.out16[c16];
                } else {
                    for (uint32_t q = (b0 > lo ? b0 : lo); q < b0 + 16u && q < hi; q++) g16[q] = sout[q];
                }
            }
        }
        __syncthreads();
    }
    st.out += emit;
    const uint32_t ctl = (m < kMax) ? (sm.tcode[m] & 0xFFFFu) : 257u;   // no control code within a full table: overflow, stop
    if (ctl != 256u) return 1;
    st.bitpos += lzw_cum_bits((uint32_t)m) + lzw_width((uint32_t)m);
    st.buf ^= 1;
    return 0;
}

__global__ void __launch_bounds__(kLzwThreads, 4)
lzw_kernel(const uint8_t* __restrict__ blob, const b2_stream_desc* __restrict__ streams, const int* __restrict__ order,
           int n_streams, uint8_t* scratch, int32_t* __restrict__ status, unsigned int* next_stream) {
    __shared__ LzwSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {                                                           // persistent CTAs draw streams from a counter
    __syncthreads();
    if (tid == 0) sm.stream = (int)atomicAdd(next_stream, 1u);
    __syncthreads();
    // two sweeps over the stream list: the large streams (image tiles) first, the small ones (label tiles, strips) fill
    // the tail, so that the last CTAs do not finish a 512 KiB tile alone
    const int draw = sm.stream;
    if (draw >= 2 * n_streams) return;
    const int wi = draw < n_streams ? draw : draw - n_streams;
    const b2_stream_desc sd = streams[order ? order[wi] : wi];
    if (sd.codec != CODEC_LZW || (sd.dst_len >= (1u << 17)) != (draw < n_streams)) continue;
    const uint8_t* src = blob + sd.src_off;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t* const wsrc = reinterpret_cast<const uint32_t*>(src - mis);
    const uint32_t n_words = (mis + sd.src_len + 3u) >> 2;
    if (tid == 0) {
        sm.wsrc = wsrc;
        sm.dst = scratch + sd.dst_off;
        sm.n_words = n_words;
        sm.mis = mis;
        sm.src_bits = sd.src_len * 8u;
        sm.dst_len = sd.dst_len;
    }
    const uint32_t dst_len = sd.dst_len;
    LzwStream st;
    st.bitpos = st.out = 0;
    st.err = st.buf = 0;
    st.next_short = true;
    st.staged = kLzwShortWords;
    lzw_stage(sm.in[0], wsrc, (8u * mis) >> 5, n_words, tid, 0, st.staged);
    __syncthreads();                                                   // the stream's parameters are in shared memory
    while (st.out < dst_len) {
        int r = 2;
        if (st.next_short) r = lzw_pass<kLzwPerShort>(sm, st, tid, lane, warp);
        if (r == 2) {
            if (st.staged < kLzwInWords) {                             // the rest of a full segment's input
                lzw_stage(sm.in[st.buf], wsrc, (8u * mis + st.bitpos) >> 5, n_words, tid, st.staged, kLzwInWords);
                st.staged = kLzwInWords;
            }
            r = lzw_pass<kLzwPer>(sm, st, tid, lane, warp);
        }
        if (r) break;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");                   // nothing may still be landing when the next stream starts
    if (st.err == 0 && st.out < dst_len) st.err = 1;                   // stream ended early / table overflow
    if (st.err && tid == 0) set_status(status, sd.image, 10 + st.err);
  }
}

// ================================================================================================ DEFLATE
// The decode loop is one dependent chain per symbol (peek -> table -> drop), so everything on that chain is kept
// short: the compressed bytes sit in a 512-byte shared-memory window that the warp refills with coalesced loads
// one half ahead of the reader (the bit buffer is topped up with one aligned 32-bit word whose fetch was issued a
// refill earlier), and literal/length and distance codes resolve through two-level tables whose entries carry the
// base value and the number of extra bits, so no second lookup follows a hit.
constexpr int kInfWarps = 1;        // one warp per CTA: every shared-memory address in the symbol loop is an immediate
constexpr int kLitRoot = 10, kDistRoot = 8;
constexpr int kLitSub = 320, kDistSub = 256;     // second-level entries; a code set needing more decodes bit by bit
constexpr int kWinWords = 128;
constexpr int kInfStage = 256;
// chunk-parallel decode: 32 lanes x kParSub bits per chunk.  160 bits = 5 words, an odd stride: lanes in lock-step hit 32
// different banks.  The per-warp shared memory (13.9 KB) keeps 16 streams resident per SM, which a 1024-pair batch needs
// (13.8 streams per SM): larger stages / match lists cost a second wave (measured: 5.5 ms instead of 3.65 ms per batch)
constexpr int kParSub = 160;
constexpr int kParWords = 32 * kParSub / 32 + 8;     // chunk + alignment + the longest symbol (48 bits) + the words fetched ahead
constexpr int kParPre = 6;                           // words per lane requested ahead for the next chunk
constexpr int kParMatches = 128;                     // > the most matches one sub-chunk can hold (2 bits each: 80)
constexpr int kParStage = 2048;

// table entry: [15:0] value | [23:16] code bits to drop | [27:24] extra bits (BASE) or index bits (SUB) | [31:28] kind.
// One flag bit per kind makes each test in the symbol loop a single predicate-setting instruction; a literal sits in
// the low byte, where a byte store takes it without a shift; no flag = no such code.  E_SUB with value 0xFFFF marks
// a prefix whose second-level table did not fit (decode bit by bit).
enum : uint32_t { E_LIT = 1u << 28, E_BASE = 1u << 29, E_EOB = 1u << 30, E_SUB = 1u << 31 };
constexpr uint32_t kNoSub = 0xFFFFu;
__device__ __forceinline__ uint32_t ent_drop(uint32_t e) { return __byte_perm(e, 0u, 0x4442); }
__device__ __forceinline__ uint32_t ent_extra(uint32_t e) { return (e >> 24) & 15u; }
__device__ __forceinline__ uint32_t ent_value(uint32_t e) { return e & 0xFFFFu; }

struct alignas(16) InfWarpSmem {
    uint32_t win[kWinWords];
    uint32_t lit_tab[(1 << kLitRoot) + kLitSub];
    uint32_t dist_tab[(1 << kDistRoot) + kDistSub];
    uint16_t code[288];       // build scratch: rank, then bit-reversed canonical code of each symbol
    uint16_t first[16], offs[16];
    uint16_t lit_count[16], dist_count[16], cl_count[16];
    uint16_t lit_sym[288], dist_sym[32], cl_sym[19];
    uint16_t cl_lut[128];     // (symbol << 4) | length
    uint8_t lens[320];
    uint8_t cl_lens[19];
    alignas(16) uint8_t stage[kInfStage];   // literals waiting for one coalesced store
    // chunk-parallel symbol decode (inflate_par_chunk)
    uint32_t pwin[kParWords];               // the chunk's input words
    uint2 mlist[kParMatches];               // matches of the chunk in stream order: {output offset in the chunk, len | dist << 16}
    alignas(16) uint8_t pstage[kParStage];  // the chunk's literals on their way to one coalesced store
};

// Shared-memory load through a 32-bit shared-space address held in a register.  The symbol loop uses this instead of
// pointers: with pointers the compiler re-derives the shared window base (an S2R) inside the loop, on the critical path.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// the shared-space address of p as a value the compiler cannot re-derive (so it stays in its register)
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}

struct InfReader {
    const uint8_t* base;      // 16-byte aligned address at or below the first byte of the stream
    uint32_t end;             // offset from base of the first byte past the stream
    uint32_t* win;            // words [W, W+128) of the stream, W = wpos & ~63
    uint32_t win_s;           // the same, as a shared-space address (kept in a register: see lds32)
    uint32_t wpos;            // next word (from base) to append to buf
    uint32_t nextw;           // that word, fetched ahead
    uint2 nxt;                // this lane's two words of [W+128, W+192), in flight from global memory
    uint64_t buf;
    int cnt;
    int lane;

    // this lane's 8 bytes of the 256-byte half window starting at word `w`; bytes past the stream read as zero
    __device__ __forceinline__ uint2 fetch(uint32_t w) const {
        const uint64_t b = (uint64_t)w * 4 + (uint32_t)lane * 8;
        uint2 v = make_uint2(0u, 0u);
        if (b < end) {
            v = *reinterpret_cast<const uint2*>(base + b);
            if (b + 8 > end) {
                uint64_t x = ((uint64_t)v.y << 32) | v.x;
                x &= (1ull << (8 * (uint32_t)(end - b))) - 1;
                v = make_uint2((uint32_t)x, (uint32_t)(x >> 32));
            }
        }
        return v;
    }
    __device__ __forceinline__ void advance() {
        __syncwarp();
        *reinterpret_cast<uint2*>(win + (((wpos + 64) & 127) + lane * 2)) = nxt;
        nxt = fetch(wpos + 128);
        __syncwarp();
    }
    __device__ __forceinline__ void refill() {
        if (cnt <= 32) {
            buf |= (uint64_t)nextw << cnt;
            cnt += 32;
            wpos++;
            if (__builtin_expect((wpos & 63) == 0, 0)) advance();
            nextw = lds32(win_s + ((wpos & 127u) << 2));
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { buf >>= n; cnt -= n; }
    __device__ __forceinline__ uint32_t take(int n) { const uint32_t v = peek(n); drop(n); return v; }
    // offset from base of the next unconsumed byte (call on a byte boundary)
    __device__ __forceinline__ uint32_t byte_pos() const { return (uint32_t)(((uint64_t)wpos * 32 - (uint64_t)cnt) >> 3); }
    // true once more bits were consumed than the stream holds
    __device__ __forceinline__ bool overrun() const { return (int64_t)wpos * 32 - cnt > (int64_t)end * 8; }
    // bits consumed so far, counted from base
    __device__ __forceinline__ uint32_t bit_pos() const { return wpos * 32u - (uint32_t)cnt; }
    __device__ __forceinline__ void seek_bit(uint32_t bitpos) {
        seek(bitpos >> 3);
        drop((int)(bitpos & 7u));
        refill();
    }
    __device__ __forceinline__ void seek(uint32_t bytepos) {
        const uint32_t w = bytepos >> 2, W = w & ~63u;
        __syncwarp();
        const uint2 a = fetch(W), b = fetch(W + 64);
        nxt = fetch(W + 128);
        *reinterpret_cast<uint2*>(win + ((W & 127) + lane * 2)) = a;
        *reinterpret_cast<uint2*>(win + (((W + 64) & 127) + lane * 2)) = b;
        __syncwarp();
        wpos = w;
        buf = 0;
        cnt = 0;
        nextw = lds32(win_s + ((wpos & 127u) << 2));
        refill();
        drop((int)(bytepos & 3u) * 8);
        refill();
    }
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

template <bool kDist>
__device__ __forceinline__ uint32_t sym_entry(int s, int drop) {
    const uint32_t d = (uint32_t)drop << 16;
    if (kDist) return s < 30 ? (uint32_t)c_dist_base[s] | E_BASE | ((uint32_t)c_dist_extra[s] << 24) | d : d;
    if (s < 256) return (uint32_t)s | E_LIT | d;
    if (s == 256) return E_EOB | d;
    if (s < 286) return (uint32_t)c_len_base[s - 257] | E_BASE | ((uint32_t)c_len_extra[s - 257] << 24) | d;
    return d;
}

// Canonical Huffman code from lens[0..n): counts, per-length first code, symbols sorted by (length, value) and each
// symbol's bit-reversed code.  Warp-cooperative; false for a set zlib would reject (over-subscribed or incomplete).
__device__ bool canonical_codes(InfWarpSmem* sm, const uint8_t* lens, int n, uint16_t* count, uint16_t* syms, int lane, bool strict) {
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    // rank of each symbol among the earlier symbols of its length: 32 symbols per step
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int s = c0 + lane;
        const int l = s < n ? lens[s] : 0;
        const uint32_t m = __match_any_sync(0xffffffffu, l);
        const uint32_t before = count[l];
        __syncwarp();
        if (s < n) sm->code[s] = (uint16_t)(before + __popc(m & ((1u << lane) - 1u)));
        if (l && lane == __ffs(m) - 1) count[l] = (uint16_t)(before + __popc(m));
        __syncwarp();
    }
    int left = 1, longest = 0;
    uint32_t code = 0, off = 0;
    bool ok = true;
    for (int l = 1; l < 16; l++) {
        const int c = count[l];
        left = (left << 1) - c;
        if (left < 0) ok = false;
        if (c) longest = l;
        if (lane == 0) { sm->first[l] = (uint16_t)code; sm->offs[l] = (uint16_t)off; }
        code = (code + c) << 1;
        off += c;
    }
    __syncwarp();
    // zlib's inflate_table: over-subscribed sets are invalid, and so are incomplete ones unless the set is empty or
    // a single one-bit code (and never for the code-length code)
    if (!ok || (left > 0 && longest != 0 && (strict || longest != 1))) return false;
    for (int s = lane; s < n; s += 32) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t rank = sm->code[s];
        syms[sm->offs[l] + rank] = (uint16_t)s;
        sm->code[s] = (uint16_t)(__brev((uint32_t)sm->first[l] + rank) >> (32 - l));   // as read LSB-first from the stream
    }
    __syncwarp();
    return true;
}

// One-level LUT for the code-length code: (symbol << 4) | length.
__device__ bool build_cl_lut(InfWarpSmem* sm, int lane) {
    for (int i = lane; i < 128; i += 32) sm->cl_lut[i] = 0;
    if (!canonical_codes(sm, sm->cl_lens, 19, sm->cl_count, sm->cl_sym, lane, true)) return false;
    if (lane < 19) {
        const int l = sm->cl_lens[lane];
        if (l) for (uint32_t i = sm->code[lane]; i < 128u; i += (1u << l)) sm->cl_lut[i] = (uint16_t)((lane << 4) | l);
    }
    __syncwarp();
    return true;
}

// Two-level table: `root` index bits, then per-prefix second-level tables sized for the longest code under the prefix.
template <bool kDist>
__device__ bool build_table(InfWarpSmem* sm, const uint8_t* lens, int n, uint32_t* tab, int root, int sub_cap,
                            uint16_t* count, uint16_t* syms, int lane) {
    const int nroot = 1 << root;
    for (int i = lane; i < nroot + sub_cap; i += 32) tab[i] = 0;
    if (!canonical_codes(sm, lens, n, count, syms, lane, false)) return false;
    bool any_long = false;
    for (int s = lane; s < n; s += 32) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = sm->code[s];
        if (l <= root) {
            const uint32_t e = sym_entry<kDist>(s, l);
            for (uint32_t i = r; i < (uint32_t)nroot; i += (1u << l)) tab[i] = e;
        } else {
            atomicMax(tab + (r & (nroot - 1)), (uint32_t)l);       // mark: longest code under this prefix
            any_long = true;
        }
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, any_long)) return true;
    // allocate the second-level tables in prefix order
    const int per = nroot / 32;
    uint32_t mine = 0;
    for (int i = 0; i < per; i++) {
        const uint32_t e = tab[lane * per + i];
        if (e && e < 16u) mine += 1u << (e - root);
    }
    uint32_t incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    uint32_t off = incl - mine;
    for (int i = 0; i < per; i++) {
        const uint32_t e = tab[lane * per + i];
        if (e && e < 16u) {
            const uint32_t sb = e - root, size = 1u << sb;
            tab[lane * per + i] = off + size <= (uint32_t)sub_cap ? (uint32_t)(nroot + off) | E_SUB | (sb << 24) | ((uint32_t)root << 16)
                                                                  : (E_SUB | kNoSub);
            off += size;
        }
    }
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const int l = lens[s];
        if (l <= root) continue;
        const uint32_t r = sm->code[s];
        const uint32_t pe = tab[r & (nroot - 1)];
        if (!(pe & E_SUB) || ent_value(pe) == kNoSub) continue;
        const uint32_t sb = ent_extra(pe), base = ent_value(pe);
        const uint32_t e = sym_entry<kDist>(s, l - root);
        for (uint32_t i = r >> root; i < (1u << sb); i += (1u << (l - root))) tab[base + i] = e;
    }
    __syncwarp();
    return true;
}

// canonical bit-by-bit decode (puff-style) of the code at the bottom of `bits`: the fallback for prefixes whose
// second-level table did not fit.  Returns (symbol << 4) | code length, or -1.  Works on a copy of the bit buffer so
// that the reader's state never has its address taken (it must stay in registers).
__device__ __noinline__ int decode_slow(uint32_t bits, const uint16_t* count, const uint16_t* syms) {
    int code = 0, first = 0, index = 0;
    for (int l = 1; l < 16; l++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[l];
        if (code - c < first) return ((int)syms[index + (code - first)] << 4) | l;
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Second step for an E_SUB first-level entry (its drop field already applied): the symbol's own entry, its
// remaining code bits dropped.
template <bool kDist>
__device__ __forceinline__ uint32_t resolve_entry(InfReader& br, uint32_t e, uint32_t tab_s, const uint16_t* count, const uint16_t* syms) {
    if (ent_value(e) != kNoSub) {
        e = lds32(tab_s + ((ent_value(e) + br.peek(ent_extra(e))) << 2));
        br.drop(ent_drop(e));
        return e;
    }
    const int s = decode_slow((uint32_t)br.buf, count, syms);
    if (s < 0) return 0u;
    br.drop(s & 15);
    return sym_entry<kDist>(s >> 4, 0);
}

// ---- chunk-parallel symbol decode ---------------------------------------------------------------------------------
// A Huffman stream is serial only because a symbol's position is known once its predecessor is decoded — but Huffman codes
// re-synchronise: a decoder started at a wrong bit falls into step with the true symbol sequence after a few symbols.  So
// the 32 lanes each start at their own 160-bit sub-chunk of the block (lane 0 at the true position), decode to the end of
// their sub-chunk and report where the next symbol would start; every lane then restarts from where its predecessor ended
// until nothing changes any more — a fixed point that is the true decode, reached after one or two rounds when the guesses
// re-synchronise, after at most 32 when they never do (the serial order).  A last pass over the now exactly known
// symbols writes the literals (through shared memory, one coalesced store) and lists the matches, which are then copied
// lane-parallel in dependency order.  Per symbol this costs a few warp instructions instead of a warp-wide dependent
// chain; it pays for literal-dominated blocks (image data), so a block with many matches per chunk is handed to the
// serial loop.
struct ParTables {
    uint32_t lit_s, dist_s, win_s;
    const uint16_t *lit_count, *lit_sym, *dist_count, *dist_sym;
};

// 32 stream bits from bit p of the staged window
__device__ __forceinline__ uint32_t par_fetch(uint32_t win_s, uint32_t p) {
    const uint32_t a = win_s + ((p >> 5) << 2);
    return __funnelshift_r(lds32(a), lds32(a + 4), p & 31u);
}

// the code at the bottom of w: its final table entry and its length in bits (0 entry: no such code)
template <bool kDist>
__device__ __forceinline__ uint32_t par_lookup(uint32_t w, uint32_t tab_s, const uint16_t* count, const uint16_t* syms, uint32_t& nbits) {
    constexpr uint32_t root = kDist ? kDistRoot : kLitRoot;
    uint32_t e = lds32(tab_s + ((w & ((1u << root) - 1u)) << 2));
    uint32_t n = ent_drop(e);
    if (e & E_SUB) {
        if (ent_value(e) != kNoSub) {
            e = lds32(tab_s + ((ent_value(e) + ((w >> n) & ((1u << ent_extra(e)) - 1u))) << 2));
            n += ent_drop(e);
        } else {
            const int sl = decode_slow(w, count, syms);
            e = sl < 0 ? 0u : sym_entry<kDist>(sl >> 4, 0);
            n = sl < 0 ? 0u : (uint32_t)(sl & 15);
        }
    }
    nbits = n;
    return e;
}

enum : uint32_t { PAR_EOB = 1u, PAR_BAD_LIT = 2u, PAR_BAD_DIST = 4u };

// Decode the symbols that start in [start, limit) (window bit positions).  kEmit: also write the literals to `lit` (+ offset
// in the chunk's output) and append the matches to `ml`.  Returns where the next symbol starts.  The lane keeps its own
// 64-bit bit buffer over the staged window, topped up from a word fetched one refill ahead, so that the dependent chain
// of a literal is one table lookup plus a handful of ALU operations.
template <bool kEmit>
__device__ __forceinline__ uint32_t par_scan(const ParTables& t, uint32_t start, uint32_t limit, uint32_t& nout, uint32_t& nmatch,
                                             uint32_t& flags, uint8_t* lit, uint2* ml) {
    uint32_t p = start, no = nout, nm = nmatch, fl = 0;
    uint32_t wa = t.win_s + ((start >> 5) << 2);                      // shared address of the next word to append
    uint64_t buf = (((uint64_t)lds32(wa + 4) << 32) | lds32(wa)) >> (start & 31u);
    int cnt = 64 - (int)(start & 31u);
    uint32_t nextw = lds32(wa + 8);
    wa += 12;
#define B2_PAR_REFILL()                                   \
    if (cnt <= 32) {                                      \
        buf |= (uint64_t)nextw << cnt;                    \
        cnt += 32;                                        \
        nextw = lds32(wa);                                \
        wa += 4;                                          \
    }
    for (;;) {
        // literals, the bulk of an image stream, in a loop of their own: the lanes leave it one by one at their first
        // length code and the longer path below runs once for all of them, not once per symbol that any lane has there
        uint32_t w = 0, nb = 0, e = 0;
        bool more = false;
        while (p < limit) {
            B2_PAR_REFILL();
            w = (uint32_t)buf;
            e = par_lookup<false>(w, t.lit_s, t.lit_count, t.lit_sym, nb);
            if (!(e & E_LIT)) { more = true; break; }
            if (kEmit) lit[no] = (uint8_t)e;
            no++;
            p += nb;
            buf >>= nb;
            cnt -= (int)nb;
        }
        if (!more) break;
        if (e & E_EOB) { p += nb; fl = PAR_EOB; break; }
        if (!(e & E_BASE)) { fl = PAR_BAD_LIT; break; }
        const uint32_t mlen = ent_value(e) + ((w >> nb) & ((1u << ent_extra(e)) - 1u));
        nb += ent_extra(e);
        buf >>= nb;
        cnt -= (int)nb;
        B2_PAR_REFILL();
        w = (uint32_t)buf;
        uint32_t nb2;
        const uint32_t d = par_lookup<true>(w, t.dist_s, t.dist_count, t.dist_sym, nb2);
        if (!(d & E_BASE)) { fl = PAR_BAD_DIST; break; }
        const uint32_t dist = ent_value(d) + ((w >> nb2) & ((1u << ent_extra(d)) - 1u));
        nb2 += ent_extra(d);
        buf >>= nb2;
        cnt -= (int)nb2;
        if (kEmit) ml[nm] = make_uint2(no, mlen | (dist << 16));
        nm++;
        no += mlen;
        p += nb + nb2;
    }
#undef B2_PAR_REFILL
    nout = no;
    nmatch = nm;
    flags = fl;
    return p;
}

// One chunk of the current block starting at stream bit `bp` (from br.base).  Advances bp and out; sets eob when the block's
// end-of-block code was reached, dense when the chunk held so many matches that the serial loop should take over, and
// returns an inflate_kernel error code (0 = fine).
__device__ __forceinline__ int inflate_par_chunk(InfWarpSmem* sm, const ParTables& t, const uint8_t* base, uint32_t end_bytes,
                                                 uint32_t& bp, uint8_t* dst, uint32_t dst_len, uint32_t& out, bool& eob, bool& dense,
                                                 uint32_t (&pre)[kParPre], uint32_t& pre_w0, int lane) {
    constexpr uint32_t kFull = 0xffffffffu;
    // ---- stage the chunk's input: words [w0, w0 + kParWords), zeros past the stream
    // (the words of a chunk that follows a full one were requested while that one was being decoded: `pre`)
    const uint32_t w0 = bp >> 5, r0 = bp & 31u;
    const uint32_t n_words = (end_bytes + 3u) >> 2;
    const uint32_t* __restrict__ wsrc = reinterpret_cast<const uint32_t*>(base);
    __syncwarp();
    if (w0 >= pre_w0 && w0 + (uint32_t)kParWords <= pre_w0 + 32u * kParPre) {
        const uint32_t delta = w0 - pre_w0;
#pragma unroll
        for (int q = 0; q < kParPre; q++) {
            const uint32_t idx = (uint32_t)(lane + 32 * q) - delta;
            if (idx < (uint32_t)kParWords) sm->pwin[idx] = pre[q];
        }
    } else {
        for (int j = lane; j < kParWords; j += 32) {
            const uint32_t w = w0 + (uint32_t)j;
            sm->pwin[j] = w < n_words ? __ldg(wsrc + w) : 0u;
        }
    }
    pre_w0 = w0 + (uint32_t)(32 * kParSub / 32 - 4);                  // the next chunk starts 0 .. 48 bits past this one's last sub-chunk
#pragma unroll
    for (int q = 0; q < kParPre; q++) {
        const uint32_t w = pre_w0 + (uint32_t)(lane + 32 * q);
        pre[q] = w < n_words ? __ldg(wsrc + w) : 0u;
    }
    __syncwarp();
    // ---- speculative decode, then restart from the predecessor's end until the chain is consistent
    const uint32_t limit = r0 + (uint32_t)kParSub * (uint32_t)(lane + 1);
    uint32_t start = r0 + (uint32_t)kParSub * (uint32_t)lane;
    uint32_t nout = 0, nmatch = 0, flags = 0;
    uint32_t endp = par_scan<false>(t, start, limit, nout, nmatch, flags, nullptr, nullptr);
    int first_stop = 32;
    for (int round = 0; round < 34; round++) {
        const uint32_t pend = __shfl_up_sync(kFull, endp, 1);
        const uint32_t stopmask = __ballot_sync(kFull, flags != 0u);
        first_stop = stopmask ? __ffs(stopmask) - 1 : 32;
        const bool need = lane > 0 && lane <= first_stop && pend != start;
        if (!__any_sync(kFull, need)) break;
        if (need) {
            start = pend;
            nout = 0;
            nmatch = 0;
            endp = par_scan<false>(t, start, limit, nout, nmatch, flags, nullptr, nullptr);
        }
    }
    int valid = first_stop < 32 ? first_stop + 1 : 32;                // lanes on the true path
    // ---- offsets; the match list holds kParMatches entries: keep as many leading lanes as fit
    uint32_t mo = lane < valid ? nmatch : 0u, oo = lane < valid ? nout : 0u;
    uint32_t mincl = mo, oincl = oo;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(kFull, mincl, d), b = __shfl_up_sync(kFull, oincl, d);
        if (lane >= d) { mincl += a; oincl += b; }
    }
    const uint32_t fits = __ballot_sync(kFull, lane < valid && mincl <= (uint32_t)kParMatches);
    const int keep = fits == kFull ? 32 : __ffs(~fits) - 1;            // a sub-chunk holds at most kParSub / 2 matches: keep >= 1
    bool stopped = first_stop < 32;
    if (keep < valid) { valid = keep; stopped = false; }
    const uint32_t total_out = __shfl_sync(kFull, oincl, valid - 1), total_m = __shfl_sync(kFull, mincl, valid - 1);
    const uint32_t stop_flags = __shfl_sync(kFull, flags, valid - 1);
    const uint32_t chunk_end = __shfl_sync(kFull, endp, valid - 1);
    if (stopped && (stop_flags & PAR_BAD_LIT)) return 14;
    if (stopped && (stop_flags & PAR_BAD_DIST)) return 16;
    if ((uint64_t)w0 * 32u + chunk_end > (uint64_t)end_bytes * 8u) return 18;       // consumed more bits than the stream holds
    if (total_out > dst_len - out) return 3;
    // ---- emit: literals to the stage (or straight to their place when the chunk's output is larger), matches to the list
    const bool staged = total_out <= (uint32_t)kParStage;
    if (lane < valid) {
        uint32_t no = oincl - oo, nm = mincl - mo, fl;
        par_scan<true>(t, start, limit, no, nm, fl, staged ? sm->pstage : dst + out, sm->mlist);
    }
    __syncwarp();
    if (staged)
        for (uint32_t i = lane; i < total_out; i += 32) dst[out + i] = sm->pstage[i];
    __syncwarp();
    // ---- matches, lane-parallel in dependency order: a match may run once its source lies wholly before the first
    // match that has not run yet (the first one itself always may: it only reads what it or earlier bytes wrote)
    for (uint32_t mb = 0; mb < total_m; mb += 32) {
        const bool have = mb + lane < total_m;
        const uint2 me = have ? sm->mlist[mb + lane] : make_uint2(0u, 0u);
        const uint32_t mlen = me.y & 0xFFFFu, dist = me.y >> 16;
        const uint32_t o = out + me.x;
        if (__any_sync(kFull, have && dist > o)) return 17;           // reaches back before the start of the output
        const uint32_t sfrom = o - dist;
        uint32_t pending = __ballot_sync(kFull, have);
        while (pending) {
            const int first = __ffs(pending) - 1;
            const uint32_t o_first = __shfl_sync(kFull, o, first);
            const bool ready = ((pending >> lane) & 1u) && (lane == first || sfrom + mlen <= o_first);
            if (ready) {
                const uint8_t* from = dst + sfrom;
                uint8_t* to = dst + o;
                if (dist >= mlen) {                                    // disjoint: four loads in flight before the stores
                    for (uint32_t k = 0; k < mlen; k += 4) {
                        uint8_t v[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) if (k + q < mlen) v[q] = from[k + q];
#pragma unroll
                        for (int q = 0; q < 4; q++) if (k + q < mlen) to[k + q] = v[q];
                    }
                } else {                                               // overlapped: the last `dist` bytes repeat
                    for (uint32_t k = 0, r = 0; k < mlen; k++) {
                        to[k] = from[r];
                        if (++r == dist) r = 0;
                    }
                }
            }
            pending &= ~__ballot_sync(kFull, ready);
            __syncwarp();
        }
    }
    out += total_out;
    bp = w0 * 32u + chunk_end;
    eob = stopped && (stop_flags & PAR_EOB);
    dense = total_m > 3u * (uint32_t)valid;
    return 0;
}

__global__ void __launch_bounds__(kInfWarps * 32)
inflate_kernel(const uint8_t* __restrict__ blob, const b2_stream_desc* __restrict__ streams, const int* __restrict__ order,
               int n_streams, uint8_t* scratch, int32_t* __restrict__ status) {
    __shared__ InfWarpSmem smem[kInfWarps];
    InfWarpSmem* sm = smem + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int wi = blockIdx.x * kInfWarps + (threadIdx.x >> 5);
    if (wi >= n_streams) return;
    const b2_stream_desc sd = streams[order ? order[wi] : wi];
    if (sd.codec != CODEC_ZLIB) return;
    uint8_t* dst = scratch + sd.dst_off;
    const uint32_t dst_len = sd.dst_len;
    InfReader br;
    {
        const uint8_t* src = blob + sd.src_off;
        const uint32_t skip = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
        br.base = src - skip;
        br.end = skip + sd.src_len;
        br.win = sm->win;
        br.win_s = smem_addr(sm->win);
        br.lane = lane;
        br.seek(skip);
    }
    const uint32_t lit_s = smem_addr(sm->lit_tab), dist_s = smem_addr(sm->dist_tab);
    int err = 0;
    uint32_t out = 0;
    // zlib header (RFC 1950)
    {
        const uint32_t cmf = br.take(8), flg = br.take(8);
        if ((cmf & 15) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) err = 1;
    }
    // Literals go to a shared-memory stage with one uniform byte store each (all lanes store the same value, so a
    // lane always reads back what it wrote itself) and leave for global memory in one coalesced burst.
    const uint32_t stage_s = smem_addr(sm->stage);
    uint32_t sp = stage_s;                // next free stage byte
#define B2_FLUSH_LITERALS()                                                         \
    {                                                                               \
        const uint32_t n_ = sp - stage_s;                                           \
        if (n_ > dst_len - out) { err = 3; break; }                                 \
        for (uint32_t i_ = lane; i_ < n_; i_ += 32) dst[out + i_] = sm->stage[i_];  \
        out += n_;                                                                  \
        sp = stage_s;                                                               \
    }
    bool last = false;
    while (!err && !last) {
        br.refill();
        last = br.take(1);
        const uint32_t type = br.take(2);
        if (type == 0) {                  // stored
            br.drop(br.cnt & 7);
            br.refill();
            const uint32_t ln = br.take(16), nln = br.take(16);
            if ((ln ^ 0xFFFFu) != nln) { err = 2; break; }
            B2_FLUSH_LITERALS();
            if (ln > dst_len - out) { err = 3; break; }
            const uint32_t start = br.byte_pos();
            if ((uint64_t)start + ln > br.end) { err = 4; break; }
            for (uint32_t i = lane; i < ln; i += 32) dst[out + i] = br.base[start + i];
            out += ln;
            br.seek(start + ln);
            continue;
        }
        if (type == 3) { err = 5; break; }
        int nlit, ndist;
        if (type == 1) {                  // fixed Huffman
            for (int i = lane; i < 288; i += 32) sm->lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
            if (lane < 32) sm->lens[288 + lane] = 5;     // 32 five-bit codes make the set complete; 30 and 31 never decode
            nlit = 288;
            ndist = 32;
            __syncwarp();
        } else {                          // dynamic Huffman
            br.refill();
            nlit = (int)br.take(5) + 257;
            ndist = (int)br.take(5) + 1;
            const int ncl = (int)br.take(4) + 4;
            if (nlit > 286 || ndist > 30) { err = 6; break; }
            if (lane < 19) sm->cl_lens[lane] = 0;
            __syncwarp();
            for (int i = 0; i < ncl; i++) {
                br.refill();
                const uint32_t v = br.take(3);
                if (lane == 0) sm->cl_lens[c_cl_order[i]] = (uint8_t)v;
            }
            __syncwarp();
            if (!build_cl_lut(sm, lane)) { err = 7; break; }
            int idx = 0;
            int prev = 0;
            while (idx < nlit + ndist) {
                br.refill();
                const uint16_t e = sm->cl_lut[br.peek(7)];
                if (!e) { err = 8; break; }
                br.drop(e & 15);
                const int sym = e >> 4;
                int rep = 1, val = sym;
                if (sym == 16) { if (idx == 0) { err = 9; break; } val = prev; rep = 3 + (int)br.take(2); }
                else if (sym == 17) { val = 0; rep = 3 + (int)br.take(3); }
                else if (sym == 18) { val = 0; rep = 11 + (int)br.take(7); }
                if (idx + rep > nlit + ndist) { err = 10; break; }
                // lens layout: [0,nlit) literal/length, [288, 288+ndist) distance
                for (int j = lane; j < rep; j += 32) {
                    const int t = idx + j;
                    sm->lens[t < nlit ? t : 288 + (t - nlit)] = (uint8_t)val;
                }
                idx += rep;
                prev = val;
            }
            if (err) break;
            __syncwarp();
            if (sm->lens[256] == 0) { err = 11; break; }
        }
        if (!build_table<false>(sm, sm->lens, nlit, sm->lit_tab, kLitRoot, kLitSub, sm->lit_count, sm->lit_sym, lane)) { err = 12; break; }
        if (!build_table<true>(sm, sm->lens + 288, ndist, sm->dist_tab, kDistRoot, kDistSub, sm->dist_count, sm->dist_sym, lane)) { err = 13; break; }
        // ---- literal-dominated blocks: chunk-parallel decode (inflate_par_chunk); a chunk dense with matches hands
        // the rest of the block to the serial symbol loop below
        {
            B2_FLUSH_LITERALS();
            const ParTables pt{lit_s, dist_s, smem_addr(sm->pwin), sm->lit_count, sm->lit_sym, sm->dist_count, sm->dist_sym};
            uint32_t bp = br.bit_pos();
            bool eob = false, dense = false;
            uint32_t pre[kParPre], pre_w0 = 0xFFFFFFF0u;
            while (!err && !eob && !dense)
                err = inflate_par_chunk(sm, pt, br.base, br.end, bp, dst, dst_len, out, eob, dense, pre, pre_w0, lane);
            if (err) break;
            br.seek_bit(bp);
            if (eob) continue;
        }
        // ---- symbols.  After a refill the buffer holds >= 33 bits: enough for two codes (<= 15 bits each), or for
        // one length code + extra (<= 20), or one distance code + extra (<= 28).  Two first-level lookups are made
        // back to back (the second is simply discarded unless the first was a literal) and the common case, two
        // literals, costs one branch.
        while (true) {
            // Pairs of first-level literals, the bulk of an image stream, run in a PTX loop whose only taken branch
            // is its back edge (a taken branch costs a lone warp ~25 cycles, and the compiler's block layout put
            // three or four of them on this path): predicated bit-buffer top-up, two table lookups back to back
            // (the second is discarded unless both are literals), two byte stores into the stage.  It leaves with
            // why = 0: `e` is a first-level entry (its code bits dropped) that is not half of a literal pair;
            // 1: the stage is full; 2: the input window must advance before the next word can be fetched.
            uint32_t e, why;
            asm volatile(
                "{\n\t"
                ".reg .pred pneed, padv, pfull, ppair;\n\t"
                ".reg .b64 add64, b1, nw64;\n\t"
                ".reg .b32 t, idx, e1, e2, d1, d2, c1, a;\n"
                "B2_INF_LOOP:\n\t"
                "setp.gt.u32 pfull, %4, %9;\n\t"
                "@pfull bra B2_INF_FULL;\n\t"
                "setp.le.s32 pneed, %1, 32;\n\t"
                "cvt.u64.u32 nw64, %3;\n\t"
                "and.b32 t, %1, 63;\n\t"
                "shl.b64 add64, nw64, t;\n\t"
                "@pneed or.b64 %0, %0, add64;\n\t"
                "@pneed add.s32 %1, %1, 32;\n\t"
                "@pneed add.u32 %2, %2, 1;\n\t"
                "and.b32 t, %2, 63;\n\t"
                "setp.eq.and.u32 padv, t, 0, pneed;\n\t"
                "@padv bra B2_INF_ADV;\n\t"
                "and.b32 t, %2, 127;\n\t"
                "shl.b32 t, t, 2;\n\t"
                "add.u32 a, %8, t;\n\t"
                "@pneed ld.shared.u32 %3, [a];\n\t"
                "cvt.u32.u64 idx, %0;\n\t"
                "and.b32 idx, idx, 1023;\n\t"
                "shl.b32 idx, idx, 2;\n\t"
                "add.u32 a, %7, idx;\n\t"
                "ld.shared.u32 e1, [a];\n\t"
                "prmt.b32 d1, e1, 0, 0x4442;\n\t"
                "shr.u64 b1, %0, d1;\n\t"
                "sub.s32 c1, %1, d1;\n\t"
                "cvt.u32.u64 idx, b1;\n\t"
                "and.b32 idx, idx, 1023;\n\t"
                "shl.b32 idx, idx, 2;\n\t"
                "add.u32 a, %7, idx;\n\t"
                "ld.shared.u32 e2, [a];\n\t"
                "and.b32 t, e1, e2;\n\t"
                "and.b32 t, t, 0x10000000;\n\t"
                "setp.ne.u32 ppair, t, 0;\n\t"
                "@!ppair bra B2_INF_NOPAIR;\n\t"
                "st.shared.u8 [%4], e1;\n\t"
                "st.shared.u8 [%4+1], e2;\n\t"
                "add.u32 %4, %4, 2;\n\t"
                "prmt.b32 d2, e2, 0, 0x4442;\n\t"
                "shr.u64 %0, b1, d2;\n\t"
                "sub.s32 %1, c1, d2;\n\t"
                "bra B2_INF_LOOP;\n"
                "B2_INF_NOPAIR:\n\t"
                "mov.b64 %0, b1;\n\t"
                "mov.b32 %1, c1;\n\t"
                "mov.b32 %5, e1;\n\t"
                "mov.u32 %6, 0;\n\t"
                "bra B2_INF_DONE;\n"
                "B2_INF_FULL:\n\t"
                "mov.u32 %5, 0;\n\t"
                "mov.u32 %6, 1;\n\t"
                "bra B2_INF_DONE;\n"
                "B2_INF_ADV:\n\t"
                "mov.u32 %5, 0;\n\t"
                "mov.u32 %6, 2;\n"
                "B2_INF_DONE:\n\t"
                "}"
                : "+l"(br.buf), "+r"(br.cnt), "+r"(br.wpos), "+r"(br.nextw), "+r"(sp), "=r"(e), "=r"(why)
                : "r"(lit_s), "r"(br.win_s), "r"(stage_s + (uint32_t)(kInfStage - 2))
                : "memory");
            if (why == 1) {
                B2_FLUSH_LITERALS();
                continue;
            }
            if (why == 2) {
                br.advance();
                br.nextw = lds32(br.win_s + ((br.wpos & 127u) << 2));
                continue;
            }
            if (e & E_SUB) e = resolve_entry<false>(br, e, lit_s, sm->lit_count, sm->lit_sym);
            if (e & E_LIT) {
                asm volatile("st.shared.u8 [%0], %1;" ::"r"(sp), "r"(e) : "memory");
                sp += 1;
                continue;
            }
            B2_FLUSH_LITERALS();
            if (e & E_EOB) { if (br.overrun()) err = 18; break; }
            if (!(e & E_BASE)) { err = 14; break; }
            br.refill();
            const uint32_t mlen = ent_value(e) + br.take(ent_extra(e));
            br.refill();
            uint32_t d = lds32(dist_s + (br.peek(kDistRoot) << 2));
            br.drop(ent_drop(d));
            if (d & E_SUB) d = resolve_entry<true>(br, d, dist_s, sm->dist_count, sm->dist_sym);
            if (!(d & E_BASE)) { err = 16; break; }
            const uint32_t dist = ent_value(d) + br.take(ent_extra(d));
            if (dist > out) { err = 17; break; }
            if (mlen > dst_len - out) { err = 3; break; }
            // Overlapped copies repeat the last `dist` bytes periodically, so every source byte is already written:
            // issue up to four loads per lane before the first store.
            __syncwarp();
            {
                const uint8_t* from = dst + (out - dist);
                uint8_t* to = dst + out;
                for (uint32_t i0 = 0; i0 < mlen; i0 += 128) {
                    uint8_t v[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t i = i0 + k * 32 + lane;
                        if (i < mlen) v[k] = from[dist >= mlen ? i : i % dist];
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t i = i0 + k * 32 + lane;
                        if (i < mlen) to[i] = v[k];
                    }
                }
            }
            out += mlen;
            __syncwarp();
            if (br.overrun()) { err = 18; break; }
        }
    }
    while (!err && sp != stage_s) B2_FLUSH_LITERALS();
#undef B2_FLUSH_LITERALS
    if (!err && out < dst_len) err = 19;          // fewer bytes than the image needs
    if (!err) {
        // Adler-32 trailer (RFC 1950).  zlib verifies it in the inflate() call that delivers the last byte, so both
        // libpng ("IDAT: incorrect data check") and libtiff/GDAL (ZIPDecode error -> failed block read) reject a
        // stream whose check does not match: so does this decoder.
        br.drop(br.cnt & 7);
        br.refill();
        uint32_t stored = br.take(8) << 24;
        stored |= br.take(8) << 16;
        stored |= br.take(8) << 8;
        stored |= br.take(8);
        if (br.overrun()) err = 18;
        __syncwarp();
        // a = sum d_i ; b = sum ((n - i) mod 65521) d_i, in 64 bits (n < 2^32: b < 2^57).
        // 16 bytes per lane and load: a group at offset i0 adds (n - i0) S - T with S = sum d_k, T = sum k d_k
        unsigned long long a = 0, b = 0, bneg = 0;
        {
            const uint32_t head = min(out, (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
            for (uint32_t i = lane; i < head; i += 32) { const uint32_t d = dst[i]; a += d; b += (unsigned long long)((out - i) % 65521u) * d; }
            const uint32_t groups = (out - head) >> 4;
            const uint4* g = reinterpret_cast<const uint4*>(dst + head);
            for (uint32_t q = lane; q < groups; q += 32) {
                const uint4 v = g[q];
                const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
                uint32_t S = 0, T = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t b0 = wds[k] & 0xFFu, b1 = (wds[k] >> 8) & 0xFFu, b2 = (wds[k] >> 16) & 0xFFu, b3 = wds[k] >> 24;
                    S += b0 + b1 + b2 + b3;
                    T += (4 * k) * b0 + (4 * k + 1) * b1 + (4 * k + 2) * b2 + (4 * k + 3) * b3;
                }
                a += S;
                b += (unsigned long long)((out - head - (q << 4)) % 65521u) * S;
                bneg += T;
            }
            for (uint32_t i = head + (groups << 4) + lane; i < out; i += 32) { const uint32_t d = dst[i]; a += d; b += (unsigned long long)((out - i) % 65521u) * d; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            bneg += __shfl_xor_sync(0xffffffffu, bneg, o);
        }
        b = b % 65521ull + 65521ull - bneg % 65521ull;
        const uint32_t s1 = (uint32_t)((a + 1ull) % 65521ull);
        const uint32_t s2 = (uint32_t)((b + (unsigned long long)(out % 65521u)) % 65521ull);
        if (!err && ((s2 << 16) | s1) != stored) err = 20;
    }
    if (err && lane == 0) set_status(status, sd.image, 20 + err);
}

// ================================================================================================ raw copy
__global__ void __launch_bounds__(256)
rawcopy_kernel(const uint8_t* __restrict__ blob, const b2_stream_desc* __restrict__ streams, int n_streams,
               uint8_t* __restrict__ scratch, int32_t* __restrict__ status) {
    const int si = blockIdx.y;
    if (si >= n_streams) return;
    const b2_stream_desc sd = streams[si];
    if (sd.codec != CODEC_RAW) return;
    if (sd.src_len < sd.dst_len) {
        if (blockIdx.x == 0 && threadIdx.x == 0) set_status(status, sd.image, 31);
        return;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sd.dst_len; i += gridDim.x * blockDim.x)
        scratch[sd.dst_off + i] = blob[sd.src_off + i];
}

// ================================================================================================ PNG un-filter
// One warp per image.  Rows are processed in bands of 32; lane L owns row band*32+L and runs x = step - L, so
// the row above is always exactly one pixel ahead: `up` arrives by shuffle from lane L-1, `up-left` is the `up`
// of the previous step, `left` is the lane's own previous result.  A "pixel" is the PNG filter unit: the bytes of
// one complete pixel, at least one (1..8).  8-bit grey / grey+alpha / RGB / RGBA land in the output directly;
// the other flavours (palette, 1/2/4-bit, 16-bit) are un-filtered into scratch and expanded by the same warp the way
// the replaced decoder presents them (b2chips.h: B2_PNG_AS_TF or GDAL).
constexpr int kUfWarps = 8;                 // warps per image of the wide kernel
constexpr int kUfRowMax = 4096;             // longest row (bytes) it takes: the last row of a band is handed on through shared memory

// 8-bit grey / grey+alpha / RGB / RGBA, progressive, at least three bands of rows: png_unfilter_wide_kernel's share
__host__ __device__ __forceinline__ bool png_wide_eligible(const b2_image_desc& im) {
    const int depth = im.png_bit_depth ? im.png_bit_depth : 8;
    return im.format == 2 && !im.png_converted && depth == 8 && (im.png_flags & 0x100) == 0 && im.height >= 96 &&
           (size_t)im.width * im.samples <= (size_t)kUfRowMax;
}

__global__ void __launch_bounds__(128)
png_unfilter_kernel(uint8_t* __restrict__ scratch, const b2_image_desc* __restrict__ imgs, int n_images,
                    uint8_t* __restrict__ out, int32_t* __restrict__ status) {
    const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wi >= n_images) return;
    const b2_image_desc im = imgs[wi];
    if (im.format != 2) return;
    if (status && status[wi] != 0) return;
    if (png_wide_eligible(im)) return;
    const int depth = im.png_bit_depth ? im.png_bit_depth : 8, ct = im.png_color_type;
    const int src_ch = im.png_converted ? (ct == 0 || ct == 3 ? 1 : (ct == 2 ? 3 : (ct == 4 ? 2 : 4))) : im.samples;
    const int bpp = max(1, src_ch * depth / 8), h = im.height;
    const size_t rb = ((size_t)im.width * src_ch * depth + 7) / 8;
    const uint8_t* src = scratch + im.scratch_off;
    uint8_t* unf = scratch + im.scratch_off + ((im.block_bytes + 15) & ~(uint64_t)15);
    uint8_t* dst = im.png_converted ? unf : out + im.out_off;
    bool bad = false;
    // One pass for a progressive image; the seven Adam7 passes otherwise (each a reduced image, un-filtered against
    // its own previous row and written straight to its pixels' places: row stride dy rows, pixel stride dx pixels).
    const bool adam7 = (im.png_flags & 0x100) != 0;
    const int full_w = (int)(rb / bpp);
    for (int pass = 0; pass < (adam7 ? 7 : 1); pass++) {
        int x0 = 0, y0 = 0, dx = 1, dy = 1;
        if (adam7) {
            x0 = (0x0402010 >> (4 * (6 - pass))) & 15;                    // 0 4 0 2 0 1 0
            y0 = (0x0040201 >> (4 * (6 - pass))) & 15;                    // 0 0 4 0 2 0 1
            dx = (0x8844221 >> (4 * (6 - pass))) & 15;                    // 8 8 4 4 2 2 1
            dy = (0x8884422 >> (4 * (6 - pass))) & 15;                    // 8 8 8 4 4 2 2
        }
        const int w = full_w > x0 ? (full_w - x0 + dx - 1) / dx : 0;      // filter units per row of this pass
        const int ph = h > y0 ? (h - y0 + dy - 1) / dy : 0;
        if (w == 0 || ph == 0) continue;
        const size_t prb = (size_t)w * bpp;                               // bytes per row of this pass
        uint8_t* pdst = dst + (size_t)y0 * rb + (size_t)x0 * bpp;
        const size_t row_stride = (size_t)dy * rb, px_stride = (size_t)dx * bpp;
        // Every lane walks its own row, so a byte-wide access is 32 separate sectors per instruction: the filtered bytes are
        // fetched an aligned 32-bit word at a time and, for a progressive image, the results leave as 32-bit words too
        // (4x fewer L1 / L2 transactions, which is what bounds this kernel).
        const bool word_out = !adam7 && (rb & 3) == 0 && (reinterpret_cast<uintptr_t>(pdst) & 3) == 0;
        for (int band = 0; band < ph; band += 32) {
            const int row = band + lane;
            const bool live = row < ph;
            const uint8_t* srow = src + (size_t)(live ? row : 0) * (prb + 1);
            const int ft = live ? srow[0] : 0;
            if (live && ft > 4) bad = true;
            uint32_t left[8] = {0, 0, 0, 0, 0, 0, 0, 0}, upl[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cur[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const uint8_t* prow = (band > 0) ? pdst + (size_t)(band - 1) * row_stride : nullptr;  // row above the band (lane 0)
            const uintptr_t in0 = reinterpret_cast<uintptr_t>(srow) + 1;                           // first filtered byte
            uint32_t rword = 0, wword = 0;
            uint8_t* const orow = pdst + (size_t)row * row_stride;
            for (int step = 0; step < w + 31; step++) {
                const int x = step - lane;
                const bool act = live && x >= 0 && x < w;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (c >= bpp) break;
                    uint32_t up = __shfl_up_sync(0xffffffffu, cur[c], 1);     // lane L-1's pixel x (computed last step)
                    if (lane == 0) up = (prow && x >= 0 && x < w) ? prow[(size_t)x * px_stride + c] : 0;
                    if (act) {
                        const uint32_t k = (uint32_t)x * bpp + c;             // byte index in the row
                        const uintptr_t ia = in0 + k;
                        if ((ia & 3) == 0 || k == 0) rword = *reinterpret_cast<const uint32_t*>(ia & ~(uintptr_t)3);
                        const uint32_t raw = (rword >> (8 * (ia & 3))) & 0xFFu;
                        const uint32_t a = left[c], b = up, cc = upl[c];
                        uint32_t pred;
                        if (ft == 0) pred = 0;
                        else if (ft == 1) pred = a;
                        else if (ft == 2) pred = b;
                        else if (ft == 3) pred = (a + b) >> 1;
                        else {
                            const int p = (int)a + (int)b - (int)cc;
                            const int pa = abs(p - (int)a), pb = abs(p - (int)b), pc = abs(p - (int)cc);
                            pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : cc);
                        }
                        const uint32_t v = (raw + pred) & 0xFFu;
                        if (word_out) {
                            wword |= v << (8 * (k & 3));
                            if ((k & 3) == 3) {
                                *reinterpret_cast<uint32_t*>(orow + (k & ~3u)) = wword;
                                wword = 0;
                            }
                        } else {
                            orow[(size_t)x * px_stride + c] = (uint8_t)v;
                        }
                        left[c] = v;
                        upl[c] = b;
                        cur[c] = v;
                    }
                }
            }
            __syncwarp();
        }
        src += (size_t)ph * (prb + 1);
    }
    if (__ballot_sync(0xffffffffu, bad)) {
        if (lane == 0) set_status(status, wi, 41);
        return;
    }
    if (!im.png_converted) return;
    // ---- expand pass
    const bool tf = (im.png_flags & B2_PNG_AS_TF) != 0;
    const uint8_t* pal = unf + ((rb * (size_t)h + 15) & ~(size_t)15);    // 256 x RGBA (palette images)
    uint8_t* o = out + im.out_off;
    const int W = im.width, S = im.samples;
    const size_t npix = (size_t)W * h;
    if (depth == 16) {                     // big-endian samples: TF keeps the high byte, GDAL the uint16 (little-endian here)
        const size_t ns = npix * src_ch;
        for (size_t i = lane; i < ns; i += 32) {
            const size_t y = i / ((size_t)W * src_ch), k = i - y * (size_t)W * src_ch;
            const uint8_t hi = unf[y * rb + 2 * k], lo = unf[y * rb + 2 * k + 1];
            if (tf) o[i] = hi;
            else { o[2 * i] = lo; o[2 * i + 1] = hi; }
        }
        return;
    }
    const uint32_t vmask = (1u << depth) - 1u;
    const uint32_t scale = depth == 1 ? 255u : (depth == 2 ? 85u : (depth == 4 ? 17u : 1u));
    for (size_t i = lane; i < npix; i += 32) {
        const size_t y = i / W;
        const uint32_t x = (uint32_t)(i - y * W);
        const uint32_t bit = x * depth;
        const uint32_t v = (unf[y * rb + (bit >> 3)] >> (8 - depth - (bit & 7))) & vmask;     // packed MSB first
        if (ct == 3 && tf) {
            for (int c = 0; c < S; c++) o[i * S + c] = pal[4 * v + c];
        } else {
            o[i] = (uint8_t)(ct == 0 && tf ? v * scale : v);
        }
    }
}

// The same wavefront with EIGHT warps per image: warp w takes the bands of 32 rows w, w + 8, ... and runs each exactly as the
// one-warp kernel does, except that the row above a band (the last row of the band before, another warp's work) arrives
// through shared memory: the producer's lane 31 leaves every byte it finishes in its warp's row buffer and publishes,
// every eight steps, how far it has come ((band << 13) + steps done, monotonic per warp); the consumer's lane 0 needs
// pixel x at its step x, which the producer finished at step x + 31, and polls only when the value it last saw is too
// small.  Band b + 8 reuses band b's buffer: it cannot get to pixel x before band b + 1 has read it (the chain of seven
// bands in between each waits for its predecessor to pass x).  One warp per image left the SMs at 14 warps; this fills them
// and shortens an image's critical path from 8 x (w + 31) steps to (w + 31) + 7 x ~40.
__global__ void __launch_bounds__(kUfWarps * 32)
png_unfilter_wide_kernel(uint8_t* __restrict__ scratch, const b2_image_desc* __restrict__ imgs, int n_images,
                         uint8_t* __restrict__ out, int32_t* __restrict__ status) {
    __shared__ uint8_t lastrow[kUfWarps][kUfRowMax];
    __shared__ unsigned int progress[kUfWarps];
    __shared__ int any_bad;
    const int wi = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (wi >= n_images) return;
    const b2_image_desc im = imgs[wi];
    if (!png_wide_eligible(im)) return;
    if (status && status[wi] != 0) return;
    const int bpp = im.samples, h = im.height, w = im.width;
    const size_t rb = (size_t)w * bpp;
    const uint8_t* src = scratch + im.scratch_off;
    uint8_t* dst = out + im.out_off;
    if (threadIdx.x < kUfWarps) progress[threadIdx.x] = 0;
    if (threadIdx.x == 0) any_bad = 0;
    __syncthreads();
    const bool word_out = (rb & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
    const int prod = (warp + kUfWarps - 1) % kUfWarps;                     // the warp that owns the band above
    volatile unsigned int* const vprog = progress;
    volatile uint8_t* const above = lastrow[prod];
    uint8_t* const mine = lastrow[warp];
    bool bad = false;
    unsigned int seen = 0;
    for (int band = warp; band * 32 < h; band += kUfWarps) {
        const int row = band * 32 + lane;
        const bool live = row < h;
        const uint8_t* srow = src + (size_t)(live ? row : 0) * (rb + 1);
        const int ft = live ? srow[0] : 0;
        if (live && ft > 4) bad = true;
        uint32_t left[4] = {0, 0, 0, 0}, upl[4] = {0, 0, 0, 0}, cur[4] = {0, 0, 0, 0};
        // the filtered bytes are fetched an aligned 32-bit word at a time: byte k of the row is byte (a0 + k) of rowwords[]
        const uintptr_t in0 = reinterpret_cast<uintptr_t>(srow) + 1;        // first filtered byte
        const uint32_t a0 = (uint32_t)(in0 & 3);
        const uint32_t* const rowwords = reinterpret_cast<const uint32_t*>(in0 - a0);
        const bool f_sub = ft == 1, f_up = ft == 2, f_avg = ft == 3, f_paeth = ft == 4;
        uint32_t rword = 0, wword = 0;
        uint8_t* const orow = dst + (size_t)row * rb;
        const unsigned int base = (unsigned int)band << 13, need0 = ((unsigned int)(band - 1) << 13) + 32u;
        for (int step = 0; step < w + 31; step++) {
            const int x = step - lane;
            const bool act = live && x >= 0 && x < w;
            if (band > 0 && step < w) {                                     // lane 0 is about to read pixel `step` of the row above
                const unsigned int need = need0 + (unsigned int)step;
                if (seen < need) {
                    do { seen = vprog[prod]; } while (seen < need);         // (a back-off or a single polling lane measured no better)
                    __threadfence_block();
                }
            }
            const uint32_t kx = (uint32_t)x * (uint32_t)bpp;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (c >= bpp) break;
                uint32_t up = __shfl_up_sync(0xffffffffu, cur[c], 1);       // lane L-1's pixel x (computed last step)
                if (lane == 0) up = (band > 0 && act) ? above[kx + c] : 0;
                if (act) {
                    const uint32_t k = kx + c, pos = a0 + k;                // byte index in the row / in rowwords
                    if ((pos & 3) == 0 || k == 0) rword = rowwords[pos >> 2];   // (fetching words ahead of use measured slower: the pass is issue-bound)
                    const uint32_t raw = (rword >> (8 * (pos & 3))) & 0xFFu;
                    const uint32_t a = left[c], b = up, cc = upl[c];
                    // every predictor, then a select on the row's filter type (the lanes of a warp hold rows of different
                    // types: branches would run one after the other)
                    const int pa = abs((int)b - (int)cc), pb = abs((int)a - (int)cc), pc = abs((int)a + (int)b - 2 * (int)cc);
                    const uint32_t paeth = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : cc);
                    uint32_t pred = f_sub ? a : 0u;
                    pred = f_up ? b : pred;
                    pred = f_avg ? (a + b) >> 1 : pred;
                    pred = f_paeth ? paeth : pred;
                    const uint32_t v = (raw + pred) & 0xFFu;
                    if (word_out) {
                        wword |= v << (8 * (k & 3));
                        if ((k & 3) == 3) {
                            *reinterpret_cast<uint32_t*>(orow + (k & ~3u)) = wword;
                            wword = 0;
                        }
                    } else {
                        orow[k] = (uint8_t)v;
                    }
                    if (lane == 31) mine[k] = (uint8_t)v;                   // the row the next band starts from
                    left[c] = v;
                    upl[c] = b;
                    cur[c] = v;
                }
            }
            if ((step & 7) == 7 || step == w + 30) {                        // publish: `step + 1` steps of this band are done
                __syncwarp();
                if (lane == 31) {
                    __threadfence_block();
                    vprog[warp] = base + (unsigned int)step + 1u;
                }
            }
        }
    }
    if (__ballot_sync(0xffffffffu, bad) && lane == 0) atomicOr(&any_bad, 1);
    __syncthreads();
    if (threadIdx.x == 0 && any_bad) set_status(status, wi, 41);
}

// ================================================================================================ TIFF assembly
// predictor 2 undo, in place on the decoded block scratch: one warp per block row, one channel at a time,
// chunked warp scan (each lane sums a run of pixels, exclusive scan across lanes, then rewrites its run).
template <typename T>
__device__ __forceinline__ T load_sample(const uint8_t* p, bool swap) {
    T v;
    memcpy(&v, p, sizeof(T));
    if (swap) {
        if (sizeof(T) == 2) v = (T)__byte_perm((uint32_t)v, 0, 0x0001);
        else if (sizeof(T) == 4) v = (T)__byte_perm((uint32_t)v, 0, 0x0123);
    }
    return v;
}

template <typename T>
__device__ __forceinline__ T swap_sample(T v) {
    if (sizeof(T) == 2) return (T)__byte_perm((uint32_t)v, 0, 0x0001);
    if (sizeof(T) == 4) return (T)__byte_perm((uint32_t)v, 0, 0x0123);
    return v;
}

// One warp per block row.  kSpb > 0: the row is walked in coalesced 512-byte chunks (16 bytes per lane); each lane sums
// its samples per channel, one warp scan per channel turns the sums into running totals, and the lane rewrites its 16
// bytes in place.  kSpb == 0 (more than 4 interleaved samples, or rows that are not 16-byte multiples): the generic walk,
// one channel at a time with each lane owning a run of consecutive pixels.
template <typename T, int kSpb>
__global__ void __launch_bounds__(128)
hdiff_undo_kernel(uint8_t* __restrict__ scratch, const b2_image_desc* __restrict__ imgs, int img_base,
                  const int32_t* __restrict__ status) {
    const int img_index = img_base + blockIdx.y;              // one launch covers a batch of images
    const b2_image_desc im = imgs[img_index];
    if (im.format != 1 || im.predictor != 2 || im.bytes_per_sample != (int)sizeof(T)) return;
    if (status && status[img_index] != 0) return;
    const int planes = im.planar == 2 ? im.samples : 1;
    const int spb = im.planar == 2 ? 1 : im.samples;
    const bool fast_ok = spb >= 1 && spb <= 4 && (((size_t)im.block_w * spb * sizeof(T)) & 15) == 0 && (im.block_bytes & 15) == 0 &&
                         (im.scratch_off & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0;
    if (kSpb == 0 ? fast_ok : (!fast_ok || spb != kSpb)) return;        // exactly one instantiation handles an image
    const long long rows_total = (long long)im.blocks_across * im.blocks_down * planes * im.block_h;
    const long long wrow = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wrow >= rows_total) return;
    const long long blk = wrow / im.block_h;
    const int ry = (int)(wrow % im.block_h);
    uint8_t* base = scratch + im.scratch_off + (uint64_t)blk * im.block_bytes + (uint64_t)ry * im.block_w * spb * sizeof(T);
    const bool swap = im.big_endian && sizeof(T) > 1;
    const int bw = im.block_w;
    if (kSpb > 0) {
        constexpr int V = 16 / (int)sizeof(T);                            // samples per lane per chunk
        constexpr int S = kSpb > 0 ? kSpb : 1;
        const uint32_t n_samples = (uint32_t)bw * S;
        T carry[S];
#pragma unroll
        for (int c = 0; c < S; c++) carry[c] = 0;
        for (uint32_t s0 = 0; s0 < n_samples; s0 += 32 * V) {
            const uint32_t mine = s0 + lane * V;                           // first sample of this lane
            const bool live = mine < n_samples;                            // rows are multiples of 16 bytes: all or nothing
            T v[V];
            uint4 q = make_uint4(0, 0, 0, 0);
            if (live) q = *reinterpret_cast<const uint4*>(base + (size_t)mine * sizeof(T));
            memcpy(v, &q, 16);
            if (swap) {
#pragma unroll
                for (int j = 0; j < V; j++) v[j] = swap_sample<T>(v[j]);
            }
            const int ph = (int)(mine % S);                                // channel of v[0]
            T loc[S], incl[S];
#pragma unroll
            for (int c = 0; c < S; c++) loc[c] = 0;
#pragma unroll
            for (int j = 0; j < V; j++) {
                const int ch = (ph + j) % S;
#pragma unroll
                for (int c = 0; c < S; c++) if (ch == c) loc[c] += v[j];
            }
#pragma unroll
            for (int c = 0; c < S; c++) {
                T x = loc[c];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const T t = (T)__shfl_up_sync(0xffffffffu, (uint32_t)x, o);
                    if (lane >= o) x += t;
                }
                incl[c] = x;
            }
            T run[S];
#pragma unroll
            for (int c = 0; c < S; c++) run[c] = (T)(carry[c] + incl[c] - loc[c]);
#pragma unroll
            for (int j = 0; j < V; j++) {
                const int ch = (ph + j) % S;
#pragma unroll
                for (int c = 0; c < S; c++)
                    if (ch == c) { run[c] += v[j]; v[j] = run[c]; }
            }
#pragma unroll
            for (int c = 0; c < S; c++) carry[c] = (T)(carry[c] + (T)__shfl_sync(0xffffffffu, (uint32_t)incl[c], 31));
            if (live) {
                if (swap) {                                                // store back in FILE byte order; assemble swaps once
#pragma unroll
                    for (int j = 0; j < V; j++) v[j] = swap_sample<T>(v[j]);
                }
                memcpy(&q, v, 16);
                *reinterpret_cast<uint4*>(base + (size_t)mine * sizeof(T)) = q;
            }
        }
        return;
    }
    const int per = (bw + 31) / 32;
    const int x0 = lane * per, x1 = min(bw, x0 + per);
    for (int c = 0; c < spb; c++) {
        T sum = 0;
        for (int x = x0; x < x1; x++) sum += load_sample<T>(base + ((size_t)x * spb + c) * sizeof(T), swap);
        T incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T t = (T)__shfl_up_sync(0xffffffffu, (uint32_t)incl, o);
            if (lane >= o) incl += t;
        }
        T run = incl - sum;
        for (int x = x0; x < x1; x++) {
            uint8_t* p = base + ((size_t)x * spb + c) * sizeof(T);
            run += load_sample<T>(p, swap);
            T v = run;                                             // store back in FILE byte order; assemble swaps once
            if (swap) v = swap_sample<T>(v);
            memcpy(p, &v, sizeof(T));
        }
    }
}

// gather decoded blocks into the final (H,W,samples) array: crop edge blocks, interleave planes, fix byte order
__global__ void __launch_bounds__(256)
assemble_kernel(const uint8_t* __restrict__ scratch, const b2_image_desc* __restrict__ imgs, int n_images,
                uint8_t* __restrict__ out, const int32_t* __restrict__ status) {
    const int ii = blockIdx.y;
    if (ii >= n_images) return;
    const b2_image_desc im = imgs[ii];
    if (im.format != 1) return;
    if (status && status[ii] != 0) return;
    const int bs = im.bytes_per_sample;
    const int S = im.samples, Wd = im.width, Hd = im.height;
    const uint8_t* src = scratch + im.scratch_off;
    uint8_t* dst = out + im.out_off;
    const bool fast = (im.planar == 1) && !(im.big_endian && bs > 1);
    if (fast) {
        // row segments are contiguous in both layouts
        const uint64_t rowbytes = (uint64_t)Wd * S * bs;
        const uint64_t total = rowbytes * Hd;
        const uint64_t blk_rowbytes = (uint64_t)im.block_w * S * bs;
        if ((rowbytes & 15) == 0 && (blk_rowbytes & 15) == 0 && total < (1ull << 35) &&
            ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | im.block_bytes) & 15) == 0) {
            // every 16-byte group of the output lies inside one block row: one 128-bit load + store, 32-bit index math
            const uint32_t gpr = (uint32_t)(rowbytes >> 4), bgpr = (uint32_t)(blk_rowbytes >> 4);
            const uint32_t groups = gpr * (uint32_t)Hd;
            const uint32_t blk_g = (uint32_t)(im.block_bytes >> 4);
            for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += gridDim.x * blockDim.x) {
                const uint32_t y = g / gpr, xg = g - y * gpr;
                const uint32_t bx = xg / bgpr, by = y / (uint32_t)im.block_h;
                const uint64_t sg = (uint64_t)(by * (uint32_t)im.blocks_across + bx) * blk_g +
                                    (uint64_t)(y - by * (uint32_t)im.block_h) * bgpr + (xg - bx * bgpr);
                st_cs(reinterpret_cast<uint4*>(dst) + g, ld_nc(reinterpret_cast<const uint4*>(src) + sg));
            }
            return;
        }
        for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (uint64_t)gridDim.x * blockDim.x * 4) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t o = i + j;
                if (o >= total) break;
                const uint32_t y = (uint32_t)(o / rowbytes);
                const uint64_t xb = o - (uint64_t)y * rowbytes;
                const uint32_t bx = (uint32_t)(xb / blk_rowbytes), by = y / im.block_h;
                const uint64_t inb = xb - (uint64_t)bx * blk_rowbytes;
                const uint64_t so = ((uint64_t)by * im.blocks_across + bx) * im.block_bytes + (uint64_t)(y - by * im.block_h) * blk_rowbytes + inb;
                v |= (uint32_t)src[so] << (8 * j);
            }
            if (i + 4 <= total && ((reinterpret_cast<uintptr_t>(dst) + i) & 3) == 0) {
                *reinterpret_cast<uint32_t*>(dst + i) = v;
            } else {
                for (int j = 0; j < 4 && i + j < total; j++) dst[i + j] = (uint8_t)(v >> (8 * j));
            }
        }
        return;
    }
    const uint64_t n_samples = (uint64_t)Wd * Hd * S;
    const int spb = im.planar == 2 ? 1 : S;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_samples; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i % S);
        const uint64_t pix = i / S;
        const uint32_t x = (uint32_t)(pix % Wd), y = (uint32_t)(pix / Wd);
        const uint32_t bx = x / im.block_w, by = y / im.block_h;
        const uint32_t plane = im.planar == 2 ? c : 0, cc = im.planar == 2 ? 0 : c;
        const uint64_t blk = ((uint64_t)plane * im.blocks_down + by) * im.blocks_across + bx;
        const uint64_t so = blk * im.block_bytes + (((uint64_t)(y - by * im.block_h) * im.block_w + (x - bx * im.block_w)) * spb + cc) * bs;
        const uint8_t* p = src + so;
        uint8_t* q = dst + i * bs;
        if (im.big_endian) for (int j = 0; j < bs; j++) q[j] = p[bs - 1 - j];
        else for (int j = 0; j < bs; j++) q[j] = p[j];
    }
}

}  // namespace b2

using namespace b2;

extern "C" int b2_decode_streams(b2_ctx* ctx, const uint8_t* blob, const b2_stream_desc* streams, int n_streams,
                                 uint32_t codec_mask, uint32_t max_raw_len, uint8_t* scratch, int32_t* status, b2_stream stream) {
    B2_REQUIRE(ctx && blob && streams && scratch && status, "b2_decode_streams: NULL argument");
    if (n_streams <= 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (codec_mask & 1u) {   // LZW: persistent CTAs, one stream at a time each
        static int per_sm = 0;
        if (!per_sm) {
            B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lzw_kernel, kLzwThreads, 0));
            if (per_sm < 1) per_sm = 1;
        }
        unsigned ctas = (unsigned)n_streams;
        const unsigned resident = (unsigned)(ctx->sm_count * per_sm);
        if (ctas > resident) ctas = resident;
        WsLock ws_lock(ctx);
        if (int e = ws_reserve(ctx, 256, s)) return e;
        unsigned int* counter = reinterpret_cast<unsigned int*>(ctx->ws);
        B2_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
        lzw_kernel<<<ctas, kLzwThreads, 0, s>>>(blob, streams, nullptr, n_streams, scratch, status, counter);
        ctx->launches++;
        B2_CUDA(cudaGetLastError());
    }
    if (codec_mask & 2u) {   // zlib / DEFLATE
        inflate_kernel<<<(n_streams + kInfWarps - 1) / kInfWarps, kInfWarps * 32, 0, s>>>(blob, streams, nullptr, n_streams, scratch, status);
        ctx->launches++;
        B2_CUDA(cudaGetLastError());
    }
    if (codec_mask & 4u) {   // uncompressed
        unsigned gx = (max_raw_len + 256 * 16 - 1) / (256 * 16);
        if (gx < 1) gx = 1;
        if (gx > 64) gx = 64;
        for (int s0 = 0; s0 < n_streams; s0 += 65535) {
            const int m = n_streams - s0 < 65535 ? n_streams - s0 : 65535;
            rawcopy_kernel<<<dim3(gx, m), 256, 0, s>>>(blob, streams + s0, m, scratch, status);
            ctx->launches++;
        }
        B2_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int b2_assemble_images(b2_ctx* ctx, uint8_t* scratch, const b2_image_desc* imgs_dev, const b2_image_desc* imgs_host,
                                  int n_images, uint8_t* out, int32_t* status, b2_stream stream) {
    B2_REQUIRE(ctx && scratch && imgs_dev && imgs_host && out && status, "b2_assemble_images: NULL argument");
    if (n_images <= 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    bool any_png = false, any_tiff = false;
    int wide_row = 0;                               // longest row among the PNGs the eight-warp un-filter takes
    uint64_t max_bytes = 0;
    long long max_rows[5] = {0, 0, 0, 0, 0};       // predictor-2 rows of the largest image, per sample size
    unsigned spb_mask[5] = {0, 0, 0, 0, 0};        // interleave factors present (bit 0: images that need the generic walk)
    for (int i = 0; i < n_images; i++) {
        const b2_image_desc& im = imgs_host[i];
        if (im.format == 2) any_png = true;
        if (png_wide_eligible(im) && im.width * im.samples > wide_row) wide_row = im.width * im.samples;
        if (im.format == 1) {
            any_tiff = true;
            const uint64_t nb = (uint64_t)im.width * im.height * im.samples * im.bytes_per_sample;
            if (nb > max_bytes) max_bytes = nb;
            if (im.predictor == 2) {
                if (im.bytes_per_sample != 1 && im.bytes_per_sample != 2 && im.bytes_per_sample != 4)
                    return fail("b2_assemble_images: predictor 2 needs 8/16/32-bit samples");
                const int planes = im.planar == 2 ? im.samples : 1;
                const long long rows = (long long)im.blocks_across * im.blocks_down * planes * im.block_h;
                if (rows > max_rows[im.bytes_per_sample]) max_rows[im.bytes_per_sample] = rows;
                const int spb = im.planar == 2 ? 1 : im.samples;
                const bool fast = spb >= 1 && spb <= 4 && (((size_t)im.block_w * spb * im.bytes_per_sample) & 15) == 0 &&
                                  (im.block_bytes & 15) == 0 && (im.scratch_off & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0;
                spb_mask[im.bytes_per_sample] |= 1u << (fast ? spb : 0);
            }
        }
    }
    for (int bs = 1; bs <= 4; bs <<= 1) {          // one launch per sample size and 65535 images (grid.y = image)
        if (!max_rows[bs]) continue;
        const unsigned gx = (unsigned)((max_rows[bs] * 32 + 127) / 128);
        for (int s0 = 0; s0 < n_images; s0 += 65535) {
            const int m = n_images - s0 < 65535 ? n_images - s0 : 65535;
            for (int sp = 0; sp <= 4; sp++) {       // one instantiation per interleave factor present (0 = generic walk)
                if (!(spb_mask[bs] & (1u << sp))) continue;
#define B2_HDIFF(TT, SP) hdiff_undo_kernel<TT, SP><<<dim3(gx, m), 128, 0, s>>>(scratch, imgs_dev, s0, status)
#define B2_HDIFF_T(TT) (sp == 0 ? B2_HDIFF(TT, 0) : sp == 1 ? B2_HDIFF(TT, 1) : sp == 2 ? B2_HDIFF(TT, 2) : sp == 3 ? B2_HDIFF(TT, 3) : B2_HDIFF(TT, 4))
                if (bs == 1) B2_HDIFF_T(uint8_t);
                else if (bs == 2) B2_HDIFF_T(uint16_t);
                else B2_HDIFF_T(uint32_t);
#undef B2_HDIFF_T
#undef B2_HDIFF
                ctx->launches++;
            }
        }
    }
    B2_CUDA(cudaGetLastError());
    if (any_tiff) {
        unsigned gx = (unsigned)((max_bytes / 4 + 255) / 256);
        if (gx < 1) gx = 1;
        if (gx > 256) gx = 256;
        for (int s0 = 0; s0 < n_images; s0 += 65535) {
            const int m = n_images - s0 < 65535 ? n_images - s0 : 65535;
            assemble_kernel<<<dim3(gx, m), 256, 0, s>>>(scratch, imgs_dev + s0, m, out, status + s0);
            ctx->launches++;
        }
        B2_CUDA(cudaGetLastError());
    }
    if (any_png) {
        png_unfilter_kernel<<<(n_images * 32 + 127) / 128, 128, 0, s>>>(scratch, imgs_dev, n_images, out, status);   // (+ expand pass)
        ctx->launches++;
        if (wide_row) {                                                    // its share of the images
            png_unfilter_wide_kernel<<<n_images, kUfWarps * 32, 0, s>>>(scratch, imgs_dev, n_images, out, status);
            ctx->launches++;
        }
        B2_CUDA(cudaGetLastError());
    }
    return 0;
}

// ================================================================================================ host parsers
namespace {
struct Rd {
    const uint8_t* b;
    uint64_t n;
    bool be;
    bool ok(uint64_t o, uint64_t k) const { return o <= n && k <= n - o; }
    uint16_t u16(uint64_t o) const { return be ? (uint16_t)((b[o] << 8) | b[o + 1]) : (uint16_t)(b[o] | (b[o + 1] << 8)); }
    uint32_t u32(uint64_t o) const {
        return be ? ((uint32_t)b[o] << 24) | ((uint32_t)b[o + 1] << 16) | ((uint32_t)b[o + 2] << 8) | b[o + 3]
                  : (uint32_t)b[o] | ((uint32_t)b[o + 1] << 8) | ((uint32_t)b[o + 2] << 16) | ((uint32_t)b[o + 3] << 24);
    }
};
const int kTypeSize[17] = {0, 1, 1, 2, 4, 8, 1, 1, 2, 4, 8, 4, 8, 0, 0, 0, 8};

struct TiffTags {
    uint32_t width = 0, height = 0, compression = 1, spp = 1, planar = 1, predictor = 1, fill_order = 1;
    uint32_t bps = 1, fmt = 1, tile_w = 0, tile_h = 0, rps = 0;
    bool mixed = false, has_size = false;
    uint64_t off_pos = 0, off_cnt = 0, cnt_pos = 0, cnt_cnt = 0;
    int off_type = 0, cnt_type = 0;
    bool has_nodata = false;
    double nodata = 0;
    // GeoTIFF: ModelPixelScale (33550), ModelTiepoint (33922), ModelTransformation (34264), GeoKeyDirectory (34735)
    int n_scale = 0, n_tie = 0, n_xform = 0;
    double scale[3] = {0, 0, 0}, tie[6] = {0, 0, 0, 0, 0, 0}, xform[16] = {0};
    int epsg = 0;              // ProjectedCSTypeGeoKey (3072) or GeographicTypeGeoKey (2048) when it is an EPSG code
};

// value i of an IFD entry of SHORT/LONG type
uint32_t entry_val(const Rd& r, int type, uint64_t pos, uint64_t i) { return type == 3 ? r.u16(pos + 2 * i) : r.u32(pos + 4 * i); }

int parse_tiff(const uint8_t* blob, uint64_t size, TiffTags& t, Rd& r) {
    if (size < 8) return 2;
    r.b = blob; r.n = size;
    if (blob[0] == 'I' && blob[1] == 'I') r.be = false;
    else if (blob[0] == 'M' && blob[1] == 'M') r.be = true;
    else return 2;
    if (r.u16(2) != 42) return 3;                       // BigTIFF / unknown
    const uint64_t ifd = r.u32(4);
    if (!r.ok(ifd, 2)) return 2;
    const uint32_t n = r.u16(ifd);
    if (!r.ok(ifd + 2, (uint64_t)n * 12)) return 2;
    bool hw = false, hh = false;
    for (uint32_t i = 0; i < n; i++) {
        const uint64_t e = ifd + 2 + 12ull * i;
        const uint32_t tag = r.u16(e), type = r.u16(e + 2), cnt = r.u32(e + 4);
        if (type == 0 || type > 16 || kTypeSize[type] == 0) continue;
        const uint64_t nb = (uint64_t)kTypeSize[type] * cnt;
        const uint64_t pos = nb <= 4 ? e + 8 : r.u32(e + 8);
        if (!r.ok(pos, nb)) return 2;
        auto v0 = [&]() -> uint32_t { return (type == 3) ? r.u16(pos) : (type == 4 ? r.u32(pos) : (type == 1 ? blob[pos] : 0)); };
        switch (tag) {
            case 256: t.width = v0(); hw = true; break;
            case 257: t.height = v0(); hh = true; break;
            case 258:
                t.bps = v0();
                for (uint32_t k = 1; k < cnt; k++) if (entry_val(r, type, pos, k) != t.bps) t.mixed = true;
                break;
            case 259: t.compression = v0(); break;
            case 266: t.fill_order = v0(); break;
            case 277: t.spp = v0(); break;
            case 278: t.rps = v0(); break;
            case 284: t.planar = v0(); break;
            case 317: t.predictor = v0(); break;
            case 322: t.tile_w = v0(); break;
            case 323: t.tile_h = v0(); break;
            case 339:
                t.fmt = v0();
                for (uint32_t k = 1; k < cnt; k++) if (entry_val(r, type, pos, k) != t.fmt) t.mixed = true;
                break;
            case 273: case 324: t.off_pos = pos; t.off_cnt = cnt; t.off_type = type; break;
            case 279: case 325: t.cnt_pos = pos; t.cnt_cnt = cnt; t.cnt_type = type; break;
            case 33550: case 33922: case 34264:
                if (type == 12) {
                    double* dst = tag == 33550 ? t.scale : (tag == 33922 ? t.tie : t.xform);
                    const uint32_t cap = tag == 33550 ? 3 : (tag == 33922 ? 6 : 16);
                    const uint32_t m = cnt < cap ? cnt : cap;
                    for (uint32_t k = 0; k < m; k++) {
                        uint64_t bits = 0;
                        for (int j = 0; j < 8; j++) bits |= (uint64_t)blob[pos + 8 * k + (r.be ? 7 - j : j)] << (8 * j);
                        memcpy(&dst[k], &bits, 8);
                    }
                    (tag == 33550 ? t.n_scale : (tag == 33922 ? t.n_tie : t.n_xform)) = (int)m;
                }
                break;
            case 34735:
                if (type == 3 && cnt >= 4) {
                    const uint32_t nkeys = r.u16(pos + 6);
                    int proj = 0, geog = 0;
                    for (uint32_t k = 0; k < nkeys && 4 * (k + 2) <= cnt; k++) {
                        const uint64_t q = pos + 8 * (k + 1);
                        const uint32_t key = r.u16(q), loc = r.u16(q + 2), val = r.u16(q + 6);
                        if (loc != 0) continue;
                        if (key == 3072) proj = (int)val;
                        else if (key == 2048) geog = (int)val;
                    }
                    t.epsg = (proj && proj != 32767) ? proj : ((geog && geog != 32767) ? geog : 0);
                }
                break;
            case 42113: {
                std::string s(reinterpret_cast<const char*>(blob + pos), (size_t)nb);
                t.has_nodata = true;
                t.nodata = atof(s.c_str());
            } break;
            default: break;
        }
    }
    t.has_size = hw && hh;
    return 0;
}

int tiff_dtype(uint32_t bps, uint32_t fmt) {
    if (bps == 8) return fmt == 2 ? B2_I8 : (fmt == 1 ? B2_U8 : -1);
    if (bps == 16) return fmt == 1 ? B2_U16 : (fmt == 2 ? B2_I16 : -1);
    if (bps == 32) return fmt == 1 ? B2_U32 : (fmt == 2 ? B2_I32 : (fmt == 3 ? B2_F32 : -1));
    if (bps == 64) return fmt == 3 ? B2_F64 : -1;
    return -1;
}
const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
const int kAdam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};   // x0, y0, dx, dy

// CRC-32 (IEEE 802.3, the PNG chunk CRC), slice-by-8.  libpng treats a CRC mismatch in a critical chunk (IHDR, IDAT)
// as a fatal error, so tf.image.decode_png / GDAL fail on such a file and the reference skips it: so do we.
#if defined(__x86_64__)
// CRC-32 (zlib polynomial, reflected) of n bytes (n >= 64, n % 16 == 0) continuing state c: four 128-bit lanes folded 64
// bytes at a time with PCLMULQDQ, then into one lane, 128 -> 64 -> 32 bits by Barrett reduction.  Constants: x^(512+64),
// x^512, x^(128+64), x^128, x^64 mod P in the reflected domain, then mu and P (Gopal et al., "Fast CRC computation for
// generic polynomials using PCLMULQDQ").  Checked against the table version for every length in tests/.
__attribute__((target("pclmul,sse4.1")))
uint32_t crc32_clmul(uint32_t c, const uint8_t* p, uint64_t n) {
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596ll, 0x0154442bd4ll);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009ell, 0x01751997d0ll);
    const __m128i k5k0 = _mm_set_epi64x(0, 0x0163cd6124ll);
    const __m128i poly = _mm_set_epi64x(0x01f7011641ll, 0x01db710641ll);
#define B2_LD128(q) _mm_loadu_si128(reinterpret_cast<const __m128i*>(q))
#define B2_FOLD128(a, b) _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128((a), k3k4, 0x11), (b)), _mm_clmulepi64_si128((a), k3k4, 0x00))
    __m128i x1 = _mm_xor_si128(B2_LD128(p), _mm_cvtsi32_si128((int)c)), x2 = B2_LD128(p + 16), x3 = B2_LD128(p + 32), x4 = B2_LD128(p + 48);
    p += 64;
    n -= 64;
    while (n >= 64) {
        const __m128i l1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), l2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        const __m128i l3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), l4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k1k2, 0x11), l1), B2_LD128(p));
        x2 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x2, k1k2, 0x11), l2), B2_LD128(p + 16));
        x3 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x3, k1k2, 0x11), l3), B2_LD128(p + 32));
        x4 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x4, k1k2, 0x11), l4), B2_LD128(p + 48));
        p += 64;
        n -= 64;
    }
    x1 = B2_FOLD128(x1, x2);
    x1 = B2_FOLD128(x1, x3);
    x1 = B2_FOLD128(x1, x4);
    while (n >= 16) {
        x2 = B2_LD128(p);
        x1 = B2_FOLD128(x1, x2);
        p += 16;
        n -= 16;
    }
    const __m128i lo32 = _mm_setr_epi32(~0, 0, ~0, 0);
    x2 = _mm_clmulepi64_si128(x1, k3k4, 0x10);                            // 128 -> 64 bits
    x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x2);
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(x1, lo32), k5k0, 0x00), x2);
    x2 = _mm_clmulepi64_si128(_mm_and_si128(x1, lo32), poly, 0x10);      // Barrett
    x2 = _mm_clmulepi64_si128(_mm_and_si128(x2, lo32), poly, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return (uint32_t)_mm_extract_epi32(x1, 1);
}
#undef B2_LD128
#undef B2_FOLD128
#endif

struct PngCrc {
    uint32_t t[8][256];
    PngCrc() {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; i++)
            for (int s = 1; s < 8; s++) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
    }
    uint32_t update(uint32_t c, const uint8_t* p, uint64_t n) const {       // c = running state (inverted CRC)
        while (n >= 8) {
            uint32_t a, b;
            memcpy(&a, p, 4);
            memcpy(&b, p + 4, 4);
            a ^= c;
            c = t[7][a & 0xFF] ^ t[6][(a >> 8) & 0xFF] ^ t[5][(a >> 16) & 0xFF] ^ t[4][a >> 24] ^
                t[3][b & 0xFF] ^ t[2][(b >> 8) & 0xFF] ^ t[1][(b >> 16) & 0xFF] ^ t[0][b >> 24];
            p += 8;
            n -= 8;
        }
        while (n--) c = (c >> 8) ^ t[0][(c ^ *p++) & 0xFF];
        return c;
    }
    uint32_t run(const uint8_t* p, uint64_t n) const {
        uint32_t c = 0xFFFFFFFFu;
#if defined(__x86_64__)
        // the IDAT checksums are the largest item of the PNG planner's host time (1.4 GB/s per core through the tables):
        // carry-less multiplication folds 64 bytes per step (6 GB/s)
        static const bool clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
        if (clmul && n >= 64) {
            const uint64_t m = n & ~15ull;
            c = crc32_clmul(c, p, m);
            p += m;
            n -= m;
        }
#endif
        return ~update(c, p, n);
    }
};
const PngCrc kPngCrc;
// the chunk at blob + p (length n, already bounds-checked): type + data CRC against the stored one
bool png_chunk_crc_ok(const uint8_t* blob, uint64_t p, uint64_t n) {
    const uint8_t* e = blob + p + 8 + n;
    const uint32_t stored = ((uint32_t)e[0] << 24) | ((uint32_t)e[1] << 16) | ((uint32_t)e[2] << 8) | e[3];
    return kPngCrc.run(blob + p + 4, n + 4) == stored;
}
}  // namespace

// Header-only probe: what load_image_rasterio(decode=False) reads (_img_to_tf_mp.py:51-53) plus what the
// decoder needs.  info->status: 0 ok, 2 corrupt header, 3 unsupported flavour.
extern "C" int b2_image_probe(const uint8_t* blob, uint64_t size, uint32_t flags, b2_image_info* info) {
    B2_REQUIRE(blob && info, "b2_image_probe: NULL argument");
    memset(info, 0, sizeof(*info));
    if (size >= 8 && memcmp(blob, kPngSig, 8) == 0) {
        info->format = 2;
        uint64_t p = 8;
        bool ihdr = false, plte = false, trns = false, interlaced = false;
        int n_idat = 0, src_ch = 0, depth = 0, ct = 0;
        while (p + 8 <= size) {
            const uint64_t n = ((uint64_t)blob[p] << 24) | (blob[p + 1] << 16) | (blob[p + 2] << 8) | blob[p + 3];
            if (p + 12 + n > size) { info->status = 2; return 0; }
            if (memcmp(blob + p + 4, "IHDR", 4) == 0 && n >= 13) {
                const uint8_t* d = blob + p + 8;
                info->width = (int32_t)(((uint32_t)d[0] << 24) | (d[1] << 16) | (d[2] << 8) | d[3]);
                info->height = (int32_t)(((uint32_t)d[4] << 24) | (d[5] << 16) | (d[6] << 8) | d[7]);
                depth = d[8];
                ct = d[9];
                const int inter = d[12];
                src_ch = ct == 0 ? 1 : (ct == 2 ? 3 : (ct == 3 ? 1 : (ct == 4 ? 2 : (ct == 6 ? 4 : 0))));
                bool ok = src_ch != 0 && (depth == 8 || (depth == 16 && ct != 3) || ((depth == 1 || depth == 2 || depth == 4) && (ct == 0 || ct == 3)));
                if (!ok || inter > 1 || (inter == 1 && depth < 8)) info->status = 3;
                interlaced = inter == 1;
                if (!png_chunk_crc_ok(blob, p, n)) { info->status = 2; return 0; }
                ihdr = true;
            } else if (memcmp(blob + p + 4, "PLTE", 4) == 0 && n_idat == 0) {
                if (n % 3 != 0 || n > 768 || !png_chunk_crc_ok(blob, p, n)) { info->status = 2; return 0; }
                plte = true;
            } else if (memcmp(blob + p + 4, "tRNS", 4) == 0 && n_idat == 0) {
                trns = true;
            } else if (memcmp(blob + p + 4, "IDAT", 4) == 0) {
                n_idat++;
            } else if (memcmp(blob + p + 4, "IEND", 4) == 0) {
                break;
            }
            p += 12 + n;
        }
        if (!ihdr || info->width <= 0 || info->height <= 0) { info->status = 2; return 0; }
        if (info->status == 0 && ct == 3 && !plte) { info->status = 2; return 0; }
        info->png_bit_depth = depth;
        info->png_color_type = ct;
        if (flags & B2_PNG_AS_TF) {            // libpng transforms as tf.image.decode_png(dtype=uint8) sets them up
            info->samples = ct == 3 ? (trns ? 4 : 3) : src_ch;
            info->dtype = B2_U8;
        } else {                               // GDAL's PNG driver
            info->samples = src_ch;
            info->dtype = depth == 16 ? B2_U16 : B2_U8;
        }
        info->n_blocks = n_idat;          // IDAT chunks to concatenate
        info->block_w = info->width;
        info->block_h = info->height;
        info->blocks_across = info->blocks_down = 1;
        info->planar = 1;
        info->predictor = 1;
        info->compression = 8;
        const uint64_t rb = ((uint64_t)info->width * src_ch * (depth ? depth : 8) + 7) / 8;
        info->block_bytes = (uint64_t)info->height * (rb + 1);
        if (interlaced) {                      // Adam7: seven reduced images, each row with its own filter byte
            const uint64_t bpp = (uint64_t)src_ch * depth / 8;
            uint64_t total = 0;
            for (int p7 = 0; p7 < 7; p7++) {
                const int x0 = kAdam7[p7][0], y0 = kAdam7[p7][1], dx = kAdam7[p7][2], dy = kAdam7[p7][3];
                const uint64_t wp = info->width > x0 ? (uint64_t)(info->width - x0 + dx - 1) / dx : 0;
                const uint64_t hp = info->height > y0 ? (uint64_t)(info->height - y0 + dy - 1) / dy : 0;
                if (wp && hp) total += hp * (wp * bpp + 1);
            }
            info->block_bytes = total;
            info->tiled = 1;                   // PNG: marks Adam7 for the planner (the field is otherwise TIFF-only)
        }
        info->geotransform[1] = 1;
        info->geotransform[5] = 1;
        if (n_idat == 0 && info->status == 0) info->status = 2;
        return 0;
    }
    TiffTags t;
    Rd r{};
    const int pe = parse_tiff(blob, size, t, r);
    if (pe) { info->status = pe; return 0; }
    info->format = 1;
    info->width = (int32_t)t.width;
    info->height = (int32_t)t.height;
    info->samples = (int32_t)t.spp;
    info->dtype = tiff_dtype(t.bps, t.fmt);
    info->compression = (int32_t)t.compression;
    info->predictor = (int32_t)t.predictor;
    info->planar = (int32_t)t.planar;
    info->big_endian = r.be ? 1 : 0;
    info->tiled = t.tile_w ? 1 : 0;
    info->has_nodata = t.has_nodata;
    info->nodata = t.nodata;
    // affine geotransform in GDAL order (x0, dx, rx, y0, ry, dy); identity-like default when the file has none
    info->geotransform[0] = 0; info->geotransform[1] = 1; info->geotransform[2] = 0;
    info->geotransform[3] = 0; info->geotransform[4] = 0; info->geotransform[5] = 1;
    info->has_geo = 0;
    info->epsg = t.epsg;
    if (t.n_xform == 16) {
        info->geotransform[0] = t.xform[3]; info->geotransform[1] = t.xform[0]; info->geotransform[2] = t.xform[1];
        info->geotransform[3] = t.xform[7]; info->geotransform[4] = t.xform[4]; info->geotransform[5] = t.xform[5];
        info->has_geo = 1;
    } else if (t.n_scale >= 2 && t.n_tie >= 6) {
        info->geotransform[1] = t.scale[0];
        info->geotransform[5] = -t.scale[1];
        info->geotransform[0] = t.tie[3] - t.tie[0] * t.scale[0];
        info->geotransform[3] = t.tie[4] + t.tie[1] * t.scale[1];
        info->has_geo = 1;
    }
    if (!t.has_size || t.width == 0 || t.height == 0 || t.spp == 0) { info->status = 2; return 0; }
    if (info->dtype < 0 || t.mixed || t.fill_order != 1 || (t.planar != 1 && t.planar != 2) ||
        !(t.compression == 1 || t.compression == 5 || t.compression == 8 || t.compression == 32946) ||
        !(t.predictor == 1 || (t.predictor == 2 && t.bps <= 32 && t.fmt != 3))) {
        info->status = 3;
        return 0;
    }
    if (t.tile_w) { info->block_w = (int32_t)t.tile_w; info->block_h = (int32_t)t.tile_h; }
    else { info->block_w = (int32_t)t.width; info->block_h = (int32_t)((t.rps == 0 || t.rps > t.height) ? t.height : t.rps); }
    if (info->block_w <= 0 || info->block_h <= 0) { info->status = 2; return 0; }
    info->blocks_across = (info->width + info->block_w - 1) / info->block_w;
    info->blocks_down = (info->height + info->block_h - 1) / info->block_h;
    const int planes = t.planar == 2 ? (int)t.spp : 1;
    info->n_blocks = info->blocks_across * info->blocks_down * planes;
    const uint64_t spb = t.planar == 2 ? 1 : t.spp;
    info->block_bytes = (uint64_t)info->block_w * info->block_h * spb * (t.bps / 8);
    if (t.off_cnt < (uint64_t)info->n_blocks || t.cnt_cnt < (uint64_t)info->n_blocks ||
        (t.off_type != 3 && t.off_type != 4) || (t.cnt_type != 3 && t.cnt_type != 4)) { info->status = 2; return 0; }
    return 0;
}

// Offsets / byte counts of the compressed blocks (TIFF tiles or strips, in file order) or of the IDAT chunk
// payloads (PNG).  decoded_len[i] = bytes block i must produce (short last strip handled).
extern "C" int b2_image_blocks(const uint8_t* blob, uint64_t size, const b2_image_info* info, uint64_t* offsets,
                               uint64_t* counts, uint64_t* decoded_len, int cap) {
    B2_REQUIRE(blob && info && offsets && counts && decoded_len, "b2_image_blocks: NULL argument");
    B2_REQUIRE(info->status == 0, "b2_image_blocks: image was not probed successfully");
    B2_REQUIRE(cap >= info->n_blocks, "b2_image_blocks: capacity too small");
    if (info->format == 2) {
        uint64_t p = 8;
        int k = 0;
        while (p + 8 <= size) {
            const uint64_t n = ((uint64_t)blob[p] << 24) | (blob[p + 1] << 16) | (blob[p + 2] << 8) | blob[p + 3];
            if (memcmp(blob + p + 4, "IDAT", 4) == 0 && k < cap) {
                if (p + 12 + n > size || !png_chunk_crc_ok(blob, p, n)) return fail("b2_image_blocks: IDAT chunk CRC mismatch");
                offsets[k] = p + 8; counts[k] = n; decoded_len[k] = 0; k++;
            } else if (memcmp(blob + p + 4, "IEND", 4) == 0) break;
            p += 12 + n;
        }
        if (k) decoded_len[0] = info->block_bytes;
        return 0;
    }
    TiffTags t;
    Rd r{};
    if (parse_tiff(blob, size, t, r)) return fail("b2_image_blocks: corrupt TIFF");
    const uint64_t spb = t.planar == 2 ? 1 : t.spp;
    const int per_plane = info->blocks_across * info->blocks_down;
    for (int i = 0; i < info->n_blocks; i++) {
        offsets[i] = entry_val(r, t.off_type, t.off_pos, i);
        counts[i] = entry_val(r, t.cnt_type, t.cnt_pos, i);
        if (!r.ok(offsets[i], counts[i])) return fail("b2_image_blocks: block outside the file");
        uint64_t rows = info->block_h;
        if (!info->tiled) {
            const int by = (i % per_plane) / info->blocks_across;
            const uint64_t left = (uint64_t)info->height - (uint64_t)by * info->block_h;
            if (left < rows) rows = left;
        }
        decoded_len[i] = rows * info->block_w * spb * (t.bps / 8);
    }
    return 0;
}

// ================================================================================================ host: batch planner
// Plans the device work for a whole batch of encoded chips in one native call (header parse of every file, stream
// and image descriptor tables, and the gather of the compressed bytes into ONE pinned staging buffer, copied by a
// few host threads), so that the Python shim does no per-file work.  Replaces the per-file open/parse that
// rasterio / tf.io.read_file do inside the reference's worker loop (_img_to_tf_mp.py:43-53, _img_to_tf_threaded.py:87-105).
#include <thread>

namespace {
inline uint64_t up_to(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct ImgPlan {
    uint64_t stage_off, scratch_off, out_off;
    int stream0;
};

void fill_image(const uint8_t* blob, uint64_t size, int i, const b2_image_info& info, const ImgPlan& pl,
                b2_stream_desc* streams, uint8_t* stage, int32_t* status, bool inplace) {
    std::vector<uint64_t> offs(info.n_blocks), cnts(info.n_blocks), dlen(info.n_blocks);
    if (b2_image_blocks(blob, size, &info, offs.data(), cnts.data(), dlen.data(), info.n_blocks) != 0) {
        status[i] = 2;   // stream slots of this image stay codec 0 (ignored by the kernels)
        return;
    }
    if (info.format == 2) {   // PNG: concatenate the IDAT payloads into one zlib stream
        uint64_t total = 0;
        for (int k = 0; k < info.n_blocks; k++) {             // in place: the payloads only ever move towards the file's start
            memmove(stage + pl.stage_off + total, blob + offs[k], cnts[k]);
            total += cnts[k];
        }
        b2_stream_desc& sd = streams[pl.stream0];
        sd.src_off = pl.stage_off;
        sd.dst_off = pl.scratch_off;
        sd.src_len = (uint32_t)total;
        sd.dst_len = (uint32_t)info.block_bytes;
        sd.codec = CODEC_ZLIB;
        sd.image = i;
        if (info.png_color_type == 3) {   // 256-entry RGBA table from PLTE (+ tRNS): missing colours black, missing alpha opaque
            uint8_t* pal = stage + up_to(pl.stage_off + total, 16);
            for (int k = 0; k < 256; k++) { pal[4 * k] = pal[4 * k + 1] = pal[4 * k + 2] = 0; pal[4 * k + 3] = 255; }
            uint64_t p = 8;
            while (p + 8 <= size) {
                const uint64_t n = ((uint64_t)blob[p] << 24) | (blob[p + 1] << 16) | (blob[p + 2] << 8) | blob[p + 3];
                if (p + 12 + n > size) break;
                if (memcmp(blob + p + 4, "PLTE", 4) == 0) {
                    for (uint64_t k = 0; k < n / 3 && k < 256; k++) memcpy(pal + 4 * k, blob + p + 8 + 3 * k, 3);
                } else if (memcmp(blob + p + 4, "tRNS", 4) == 0) {
                    for (uint64_t k = 0; k < n && k < 256; k++) pal[4 * k + 3] = blob[p + 8 + k];
                } else if (memcmp(blob + p + 4, "IDAT", 4) == 0) {
                    break;
                }
                p += 12 + n;
            }
            b2_stream_desc& ps = streams[pl.stream0 + 1];
            ps.src_off = (uint64_t)(pal - stage);
            ps.dst_off = pl.scratch_off + up_to(info.block_bytes, 16) +
                         up_to((uint64_t)info.height * (((uint64_t)info.width * info.png_bit_depth + 7) / 8), 16);   // palette: 1 sample
            ps.src_len = ps.dst_len = 1024;
            ps.codec = CODEC_RAW;
            ps.image = i;
        }
    } else {                  // TIFF: the file as is, one stream per tile / strip
        if (!inplace) memcpy(stage + pl.stage_off, blob, size);
        const int codec = info.compression == 1 ? CODEC_RAW : (info.compression == 5 ? CODEC_LZW : CODEC_ZLIB);
        for (int k = 0; k < info.n_blocks; k++) {
            b2_stream_desc& sd = streams[pl.stream0 + k];
            sd.src_off = pl.stage_off + offs[k];
            sd.dst_off = pl.scratch_off + (uint64_t)k * info.block_bytes;
            sd.src_len = (uint32_t)cnts[k];
            sd.dst_len = (uint32_t)dlen[k];
            sd.codec = codec;
            sd.image = i;
        }
    }
}
}  // namespace

extern "C" int b2_decode_plan_batch(const uint8_t* const* blobs, const uint64_t* sizes, int n, b2_image_info* infos,
                                    int32_t* status, b2_image_desc* images, b2_stream_desc* streams, int streams_cap,
                                    uint8_t* stage, uint64_t stage_cap, int n_threads, uint32_t flags, b2_decode_plan* plan) {
    B2_REQUIRE(blobs && sizes && infos && status && images && plan, "b2_decode_plan_batch: NULL argument");
    B2_REQUIRE(n >= 0, "b2_decode_plan_batch: n < 0");
    // B2_PLAN_INPLACE: every blob already lies inside stage[0, stage_cap) (read there by b2_read_files): no gather, the
    // streams point at the files where they are (PNG IDAT payloads are compacted towards the start of their file)
    const bool inplace = (flags & B2_PLAN_INPLACE) != 0;
    flags &= ~B2_PLAN_INPLACE;
    B2_REQUIRE(!inplace || stage, "b2_decode_plan_batch: B2_PLAN_INPLACE needs the buffer that holds the files");
    memset(plan, 0, sizeof(*plan));
    std::vector<ImgPlan> pl((size_t)n);
    uint64_t stage_pos = 0, stage_hi = 0, scratch_pos = 0, out_pos = 0, compressed = 0;
    int n_streams = 0;
    uint32_t mask = 0, max_raw = 0;
    for (int i = 0; i < n; i++) {
        b2_image_info& info = infos[i];
        memset(&images[i], 0, sizeof(b2_image_desc));
        status[i] = 0;
        if (!blobs[i] || sizes[i] == 0) { memset(&info, 0, sizeof(info)); info.status = 2; status[i] = 2; continue; }
        b2_image_probe(blobs[i], sizes[i], flags, &info);
        if (info.status != 0) { status[i] = info.status; continue; }
        if (inplace) {
            B2_REQUIRE(blobs[i] >= stage && blobs[i] + sizes[i] <= stage + stage_cap,
                       "b2_decode_plan_batch: B2_PLAN_INPLACE blob outside the buffer");
            B2_REQUIRE(!(info.format == 2 && info.png_color_type == 3), "b2_decode_plan_batch: palette PNGs need the gathering mode");
            stage_pos = (uint64_t)(blobs[i] - stage);
        }
        const int bs = info.dtype == B2_U8 || info.dtype == B2_I8 ? 1 : (info.dtype == B2_U16 || info.dtype == B2_I16 ? 2 : (info.dtype == B2_F64 ? 8 : 4));
        pl[i] = ImgPlan{stage_pos, scratch_pos, out_pos, n_streams};
        b2_image_desc& im = images[i];
        im.scratch_off = scratch_pos;
        im.out_off = out_pos;
        im.block_bytes = info.block_bytes;
        im.format = info.format; im.width = info.width; im.height = info.height; im.samples = info.samples;
        im.bytes_per_sample = bs; im.predictor = info.predictor; im.planar = info.planar; im.big_endian = info.big_endian;
        im.block_w = info.block_w; im.block_h = info.block_h;
        im.blocks_across = info.blocks_across; im.blocks_down = info.blocks_down;
        if (info.format == 2) {
            im.png_bit_depth = info.png_bit_depth;
            im.png_color_type = info.png_color_type;
            im.png_flags = (int32_t)flags | (info.tiled ? 0x100 : 0);           // bit 8: Adam7
            im.png_converted = info.png_bit_depth != 8 || info.png_color_type == 3;
            n_streams += 1;
            stage_pos += sizes[i];                       // upper bound of the IDAT payload bytes
            if (im.png_converted) {                      // + un-filtered bytes + palette (see b2_image_desc)
                const int sc = info.png_color_type == 2 ? 3 : (info.png_color_type == 4 ? 2 : (info.png_color_type == 6 ? 4 : 1));
                const uint64_t unf = (uint64_t)info.height * (((uint64_t)info.width * sc * info.png_bit_depth + 7) / 8);
                scratch_pos += up_to(up_to(info.block_bytes, 16) + up_to(unf, 16) + 1024, 256);
                if (info.png_color_type == 3) {
                    n_streams += 1;                      // the palette travels as a stored stream
                    stage_pos = up_to(stage_pos, 16) + 1024;
                    mask |= 4u;
                    if (max_raw < 1024) max_raw = 1024;
                }
            } else {
                scratch_pos += up_to(info.block_bytes, 256);
            }
            mask |= 2u;
        } else {
            n_streams += info.n_blocks;
            stage_pos += sizes[i];
            scratch_pos += up_to((uint64_t)info.n_blocks * info.block_bytes, 256);
            mask |= info.compression == 1 ? 4u : (info.compression == 5 ? 1u : 2u);
            if (info.compression == 1 && info.block_bytes > max_raw) max_raw = (uint32_t)info.block_bytes;
        }
        compressed += sizes[i];
        stage_pos = up_to(stage_pos, 16);
        if (stage_pos > stage_hi) stage_hi = stage_pos;
        out_pos += up_to((uint64_t)info.width * info.height * info.samples * bs, 256);
    }
    B2_REQUIRE(!inplace || stage_hi + 16 <= stage_cap, "b2_decode_plan_batch: B2_PLAN_INPLACE needs 16 spare bytes after the last file");
    plan->stage_bytes = stage_hi + 16;
    plan->scratch_bytes = scratch_pos + 16;
    plan->out_bytes = out_pos + 16;
    plan->compressed_bytes = compressed;
    plan->n_streams = n_streams;
    plan->codec_mask = mask;
    plan->max_raw_len = max_raw;
    plan->filled = 0;
    if (!streams || !stage || streams_cap < n_streams || stage_cap < plan->stage_bytes) return 0;   // sizes only
    memset(streams, 0, sizeof(b2_stream_desc) * (size_t)n_streams);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt > 32) nt = 32;
    if (nt > n) nt = n;
    if (nt < 1) nt = 1;
    auto work = [&](int t) {
        for (int i = t; i < n; i += nt)
            if (status[i] == 0) fill_image(blobs[i], sizes[i], i, infos[i], pl[i], streams, stage, status, inplace);
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    plan->filled = 1;
    return 0;
}

// ================================================================================================ host file reader
#include <cerrno>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

extern "C" int b2_read_files(const char* const* paths, int n, uint8_t* dst, uint64_t dst_cap, uint64_t* offsets, uint64_t* sizes,
                             int32_t* status, int n_threads, uint64_t* needed) {
    B2_REQUIRE(paths && offsets && sizes && status && needed, "b2_read_files: NULL argument");
    B2_REQUIRE(n >= 0, "b2_read_files: n < 0");
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt > 32) nt = 32;
    if (nt > n) nt = n;
    if (nt < 1) nt = 1;
    auto fan = [&](auto&& body) {
        if (nt == 1) { body(0); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(body, t);
        for (auto& x : th) x.join();
    };
    fan([&](int t) {                                     // sizes
        for (int i = t; i < n; i += nt) {
            struct stat sb;
            status[i] = 0;
            sizes[i] = 0;
            if (!paths[i] || stat(paths[i], &sb) != 0) status[i] = paths[i] ? errno : EINVAL;
            else if (!S_ISREG(sb.st_mode)) status[i] = S_ISDIR(sb.st_mode) ? EISDIR : EINVAL;
            else sizes[i] = (uint64_t)sb.st_size;
        }
    });
    uint64_t pos = 0;
    for (int i = 0; i < n; i++) {
        offsets[i] = pos;
        pos += (sizes[i] + 15) & ~15ull;
    }
    *needed = pos;
    if (!dst || dst_cap < pos) return 0;
    fan([&](int t) {                                     // contents
        for (int i = t; i < n; i += nt) {
            if (status[i] != 0 || sizes[i] == 0) continue;
            const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
            if (fd < 0) { status[i] = errno; continue; }
            uint64_t got = 0;
            while (got < sizes[i]) {
                const ssize_t r = pread(fd, dst + offsets[i] + got, sizes[i] - got, (off_t)got);
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) { status[i] = r < 0 ? errno : EIO; break; }     // short file: it changed under us
                got += (uint64_t)r;
            }
            close(fd);
        }
    });
    return 0;
}
