// tfrecord.cu — K2: TFRecord framing, CRC-32C, tf.train.Example payload location, fused payload kernels.
//
// Replaces (reference call sites): tf.io.TFRecordWriter.write  _img_to_tf_mp.py:119,141 / _img_to_tf_threaded.py:182,203
//                                   Example.SerializeToString   _img_to_tf_mp.py:141, convert_to_example
//                                                               _tfrecord_image_translation.py:160-211
//                                   TFRecordDataset reader      parse_tfrecords.ipynb cell 4
//                                   parse_single_example / decode_raw / reshape
//                                                               _tfrecord_image_translation.py:249,306-314,394-407
//
// Data layout: a shard is one contiguous device buffer.  Record data is cut into 8 KiB tiles on a 16-byte
// aligned grid anchored at the record's (aligned-down) data start; one 256-thread CTA owns one tile.  A tile is
// staged ONCE in shared memory with coalesced 128-bit loads and then serves both the CRC and the payload sinks,
// so every input byte crosses HBM exactly once.
//
// CRC-32C without a CRC instruction: the pure CRC (zero init) is linear over GF(2), so
//   * thread i CRCs its two 16-byte vectors (tile offsets 16 i and 16 i + 4096) with table steps whose
//     "advance" also skips the gap between them,
//   * one GF(2)[x] multiplication by x^(8*(4080-16 i)) moves its partial to the end of the tile,
//   * partials XOR together (warp shuffles, then 8 words of shared memory) into one word per tile,
//   * a warp per record folds the tile words left to right (Horner with x^(8*8192)) and un-advances by the
//     zero padding after the record end.  The 0xFFFFFFFF init is XORed into the first four data bytes.
#include <cstring>
#include <vector>

#include "common.cuh"

namespace b2 {

struct CrcSmem {
    uint32_t t4[4][256];
    uint32_t s[4][256];
};

__device__ __forceinline__ uint32_t adv4(const uint32_t (*t)[256], uint32_t x) {
    return t[0][x & 0xff] ^ t[1][(x >> 8) & 0xff] ^ t[2][(x >> 16) & 0xff] ^ t[3][x >> 24];
}

__device__ __forceinline__ void load_crc_tables(CrcSmem* sm, const CrcTables* tab) {
    const uint32_t* g0 = &tab->t4[0][0];
    const uint32_t* g1 = &tab->s4096[0][0];
    uint32_t* d0 = &sm->t4[0][0];
    uint32_t* d1 = &sm->s[0][0];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        d0[i] = __ldg(g0 + i);
        d1[i] = __ldg(g1 + i);
    }
}

// zero the bytes of a 16-byte vector at absolute address a that fall outside [lo, hi)
__device__ __forceinline__ uint4 mask_vec(uint4 v, uint64_t a, uint64_t lo, uint64_t hi) {
    if (a >= lo && a + 16 <= hi) return v;
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t p = a + 4 * q + j;
            if (p >= lo && p < hi) m |= 0xFFu << (8 * j);
        }
        w[q] &= m;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// CRC partial of one staged tile (vectors already zero outside [d0,d1)).  Returns the CTA-wide XOR in thread 0.
// init_lo/init_hi: absolute range whose bytes get the 0xFF init XOR (d0..d0+4), only relevant for tile 0.
__device__ __forceinline__ uint32_t tile_crc(const uint4* buf4, const CrcSmem* cs, const CrcTables* tab, uint64_t ts,
                                             uint64_t d0, uint64_t d1, bool first_tile, uint32_t* red) {
    const int i = threadIdx.x;
    uint4 v0 = mask_vec(buf4[i], ts + 16ull * i, d0, d1);
    uint4 v1 = mask_vec(buf4[i + 256], ts + 4096 + 16ull * i, d0, d1);
    if (first_tile && i < 2) {  // init XOR lives in the first 4 data bytes, i.e. inside vectors 0/1 of tile 0
        uint32_t w[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t p = ts + 16ull * i + 4 * q + j;
                if (p >= d0 && p < d0 + 4) w[q] ^= 0xFFu << (8 * j);
            }
        v0 = make_uint4(w[0], w[1], w[2], w[3]);
    }
    uint32_t s = adv4(cs->t4, v0.x);
    s = adv4(cs->t4, s ^ v0.y);
    s = adv4(cs->t4, s ^ v0.z);
    s = adv4(cs->s, s ^ v0.w);
    s = adv4(cs->t4, s ^ v1.x);
    s = adv4(cs->t4, s ^ v1.y);
    s = adv4(cs->t4, s ^ v1.z);
    s = adv4(cs->t4, s ^ v1.w);
    s = multmodp(__ldg(&tab->fix[i]), s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s ^= __shfl_xor_sync(0xffffffffu, s, o);
    if ((i & 31) == 0) red[i >> 5] = s;
    __syncthreads();
    uint32_t r = 0;
    if (i == 0) {
#pragma unroll
        for (int k = 0; k < kTileThreads / 32; k++) r ^= red[k];
    }
    return r;
}

// stage [ts, ts + kTile + 32) of `base` into shared memory, zero beyond `nbytes`
__device__ __forceinline__ void stage_tile(uint4* buf4, const uint8_t* base, uint64_t ts, uint64_t nbytes) {
    for (int k = threadIdx.x; k < kTile / 16 + 2; k += blockDim.x) {
        const uint64_t a = ts + 16ull * k;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (a + 16 <= nbytes) {
            v = ld_nc(reinterpret_cast<const uint4*>(base + a));
        } else if (a < nbytes) {
            uint32_t w[4] = {0, 0, 0, 0};
            for (uint64_t p = a; p < nbytes; p++) w[(p - a) >> 2] |= (uint32_t)base[p] << (8 * ((p - a) & 3));
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        buf4[k] = v;
    }
}

__device__ __forceinline__ uint32_t smem_u32_unaligned(const uint32_t* buf32, uint32_t off) {
    const uint32_t w0 = buf32[off >> 2], w1 = buf32[(off >> 2) + 1];
    return __funnelshift_r(w0, w1, (off & 3) * 8);
}

// ---------------------------------------------------------------------------------------------- parse
struct ParseArgs {
    const uint8_t* shard;
    uint64_t nbytes;
    const uint64_t* rec_off;
    const uint64_t* rec_len;
    const b2_example_index* index;  // may be NULL (CRC-only over raw ranges)
    b2_parse_sink sink;
    const CrcTables* tab;
    uint32_t* tilecrc;  // [n][tiles_x]
    uint32_t tiles_x;
};

// raw byte copy of payload range [po, po+pl) (absolute) into dst, for the part owned by tile [ts, te)
__device__ __forceinline__ void sink_raw(const uint32_t* buf32, const uint8_t* buf8, uint64_t ts, uint64_t te,
                                         uint64_t po, uint64_t pl, uint8_t* dst) {
    if (pl == 0 || po >= te || po + pl <= ts) return;
    const bool al = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    const uint64_t lo = po > ts ? po : ts, hi = (po + pl < te) ? po + pl : te;  // absolute bytes in this tile
    if (al) {
        // 16-byte destination groups whose FIRST byte lies in the tile; last partial group done bytewise
        const uint64_t g_lo = (lo - po + 15) >> 4, g_hi = (hi - po + 15) >> 4;
        const uint64_t full = pl >> 4;
        for (uint64_t g = g_lo + threadIdx.x; g < g_hi; g += blockDim.x) {
            const uint32_t o = (uint32_t)(po + 16 * g - ts);
            if (g < full) {
                uint4 v;
                v.x = smem_u32_unaligned(buf32, o);
                v.y = smem_u32_unaligned(buf32, o + 4);
                v.z = smem_u32_unaligned(buf32, o + 8);
                v.w = smem_u32_unaligned(buf32, o + 12);
                st_cs(reinterpret_cast<uint4*>(dst + 16 * g), v);
            } else {
                for (uint64_t b = 16 * g; b < pl; b++) dst[b] = buf8[o + (b - 16 * g)];
            }
        }
        // bytes of the first (partial) group when the payload starts mid-tile are covered: g_lo*16 >= lo-po.
        // bytes between lo-po and g_lo*16 belong to a group whose first byte is in the previous tile.
    } else {
        for (uint64_t p = lo + threadIdx.x; p < hi; p += blockDim.x) dst[p - po] = buf8[p - ts];
    }
}

// Dynamic shared memory of the NORM_ONEHOT sink:
//   lut  : C*256 floats, lut[c*256+v] = (float(v) - mean[c]) / std[c] computed ONCE per CTA with IEEE division —
//          the per-pixel work is then one shared-memory lookup, and the result is bit-identical to dividing.
//   hot  : per warp, labels_per_iter*K floats kept at zero; a label sets ONE float to 1.0f, the warp streams the
//          block out with coalesced 128-bit stores and clears that float again.  No per-element compare/select.
constexpr int kLutMaxC = 8;
constexpr int kHotMaxK = 32;
__host__ __device__ inline int hot_labels_per_iter(int K) { return K <= 16 ? 64 : 32; }

template <int kMode>
__global__ void __launch_bounds__(kTileThreads)
parse_kernel(const ParseArgs a) {
    __shared__ __align__(16) uint4 buf4[kTile / 16 + 2];
    __shared__ CrcSmem cs;
    __shared__ uint32_t red[kTileThreads / 32];
    __shared__ float s_mean[64], s_std[64];
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    const int tid = threadIdx.x;
    const int r = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    const uint64_t d0 = a.rec_off[r], len = a.rec_len[r];
    const uint64_t d1 = d0 + len;
    const uint64_t A = d0 & ~15ull;
    const uint64_t ts = A + (uint64_t)tile * kTile, te = ts + kTile;
    if (len == 0 || ts >= d1) return;
    if (a.sink.verify_crc) load_crc_tables(&cs, a.tab);
    b2_example_index ix;
    bool sink_ok = false;
    const int C = a.sink.channels, K = a.sink.num_classes;
    const bool use_lut = C <= kLutMaxC, use_hot = K <= kHotMaxK;
    float* lut = reinterpret_cast<float*>(dyn_smem);
    float* hot_all = lut + (use_lut ? C * 256 : 0);
    if (kMode != B2_SINK_NONE && a.index != nullptr) {
        ix = a.index[r];
        sink_ok = ix.status == 0;
    }
    bool has_img = false, has_tgt = false;
    if (kMode == B2_SINK_NORM_ONEHOT && sink_ok) {
        sink_ok = ix.img_kind == 1 && ix.tgt_kind == 1 && ix.img_len * 4 <= a.sink.img_stride &&
                  ix.tgt_len * (uint64_t)K * 4 <= a.sink.tgt_stride && ix.img_len < (1ull << 31) &&
                  ix.tgt_len * (uint64_t)K < (1ull << 31);
        has_img = sink_ok && a.sink.img_out && ix.img_len && ix.img_off < te && ix.img_off + ix.img_len > ts;
        has_tgt = sink_ok && a.sink.tgt_out && ix.tgt_len && ix.tgt_off < te && ix.tgt_off + ix.tgt_len > ts;
        if (has_img) {
            if (use_lut) {
                for (int i = tid; i < C * 256; i += kTileThreads)
                    lut[i] = __fdiv_rn((float)(i & 255) - a.sink.mean[i >> 8], a.sink.std[i >> 8]);
            } else {
                for (int c = tid; c < C && c < 64; c += kTileThreads) {
                    s_mean[c] = a.sink.mean[c];
                    s_std[c] = a.sink.std[c];
                }
            }
        }
        if (has_tgt && use_hot) {
            const int nfl = hot_labels_per_iter(K) * K * (kTileThreads / 32);
            for (int i = tid; i < nfl; i += kTileThreads) hot_all[i] = 0.0f;
        }
    }
    stage_tile(buf4, a.shard, ts, a.nbytes < d1 ? a.nbytes : d1);
    __syncthreads();
    if (a.sink.verify_crc) {
        const uint32_t c = tile_crc(buf4, &cs, a.tab, ts, d0, d1, tile == 0, red);
        if (tid == 0) a.tilecrc[(size_t)r * a.tiles_x + tile] = c;
    }
    if (kMode == B2_SINK_NONE || !sink_ok) return;
    const uint32_t* buf32 = reinterpret_cast<const uint32_t*>(buf4);
    const uint8_t* buf8 = reinterpret_cast<const uint8_t*>(buf4);
    if (kMode == B2_SINK_RAW) {
        if (ix.img_len > a.sink.img_stride || ix.tgt_len > a.sink.tgt_stride) return;
        if (a.sink.img_out)
            sink_raw(buf32, buf8, ts, te, ix.img_off, ix.img_len, static_cast<uint8_t*>(a.sink.img_out) + (uint64_t)r * a.sink.img_stride);
        if (a.sink.tgt_out)
            sink_raw(buf32, buf8, ts, te, ix.tgt_off, ix.tgt_len, static_cast<uint8_t*>(a.sink.tgt_out) + (uint64_t)r * a.sink.tgt_stride);
        return;
    }
    // ---- NORM_ONEHOT: uint8 image -> (x-mean)/std float32 ; uint8 target -> one-hot float32
    if (has_img) {
        float* dst = reinterpret_cast<float*>(static_cast<uint8_t*>(a.sink.img_out) + (uint64_t)r * a.sink.img_stride);
        const uint64_t po = ix.img_off;
        const uint32_t pl = (uint32_t)ix.img_len;
        const uint64_t lo = po > ts ? po : ts, hi = (po + pl < te) ? po + pl : te;
        // float4 group g = image bytes [4g, 4g+4); owned by the tile that holds its first byte
        const uint32_t g_lo = (uint32_t)((lo - po + 3) >> 2), g_hi = (uint32_t)((hi - po + 3) >> 2), full = pl >> 2;
        const uint32_t base = (uint32_t)(po - ts);  // wraps when po < ts; base + 4g is back in [0, kTile)
        uint32_t g = g_lo + tid;
        uint32_t c0 = (4u * g) % (uint32_t)C;
        const uint32_t cstep = (4u * kTileThreads) % (uint32_t)C;
        for (; g < g_hi; g += kTileThreads) {
            const uint32_t x = smem_u32_unaligned(buf32, base + 4 * g);
            float f[4];
            uint32_t c = c0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t v = (x >> (8 * k)) & 0xFFu;
                f[k] = use_lut ? lut[(c << 8) + v] : __fdiv_rn((float)v - s_mean[c], s_std[c]);
                c = (c + 1 == (uint32_t)C) ? 0 : c + 1;
            }
            if (g < full) {
                st_cs(reinterpret_cast<float4*>(dst) + g, make_float4(f[0], f[1], f[2], f[3]));
            } else {
                for (uint32_t b = 4 * g; b < pl; b++) dst[b] = f[b - 4 * g];
            }
            c0 += cstep;
            if (c0 >= (uint32_t)C) c0 -= (uint32_t)C;
        }
    }
    if (has_tgt) {
        float* dst = reinterpret_cast<float*>(static_cast<uint8_t*>(a.sink.tgt_out) + (uint64_t)r * a.sink.tgt_stride);
        const uint64_t po = ix.tgt_off;
        const uint32_t pl = (uint32_t)ix.tgt_len;
        const uint64_t lo = po > ts ? po : ts, hi = (po + pl < te) ? po + pl : te;
        const uint32_t base = (uint32_t)(po - ts);
        if (use_hot) {
            // work unit = 4 labels (4K floats: always a whole number of float4s, 16-byte aligned in the output);
            // a unit belongs to the tile holding its first label, later labels may sit in the 32-byte halo
            const uint32_t j_lo = (uint32_t)((lo - po + 3) >> 2), j_hi = (uint32_t)((hi - po + 3) >> 2);
            const uint32_t L_beg = 4 * j_lo, L_end = (4 * j_hi < pl) ? 4 * j_hi : pl;
            const int Lw = hot_labels_per_iter(K);
            const int warp = tid >> 5, lane = tid & 31;
            float* hot = hot_all + warp * Lw * K;
            for (uint32_t L0 = L_beg + warp * Lw; L0 < L_end; L0 += (kTileThreads / 32) * Lw) {
                const uint32_t nl = (L_end - L0 < (uint32_t)Lw) ? L_end - L0 : (uint32_t)Lw;
                uint32_t slot[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const uint32_t li = lane + 32 * i;
                    if (li < nl) {
                        const uint32_t lab = buf8[base + L0 + li];
                        if (lab < (uint32_t)K) {
                            slot[i] = li * K + lab;
                            hot[slot[i]] = 1.0f;
                        }
                    }
                }
                __syncwarp();
                const uint32_t nfl = nl * K, nf4 = nfl >> 2;
                float4* o4 = reinterpret_cast<float4*>(dst + (size_t)L0 * K);
                const float4* h4 = reinterpret_cast<const float4*>(hot);
                for (uint32_t i = lane; i < nf4; i += 32) st_cs(o4 + i, h4[i]);
                if (lane < (nfl & 3)) dst[(size_t)L0 * K + 4 * nf4 + lane] = hot[4 * nf4 + lane];
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; i++)
                    if (slot[i] != 0xFFFFFFFFu) hot[slot[i]] = 0.0f;
            }
        } else {
            // generic path (K > 32): float4 group g holds one-hot floats [4g, 4g+4), owned by the tile of label 4g/K
            const uint32_t nfl = pl * (uint32_t)K;
            const uint32_t g_lo = (uint32_t)(((lo - po) * K + 3) >> 2), g_hi = (uint32_t)(((hi - po) * K + 3) >> 2), full = nfl >> 2;
            for (uint32_t g = g_lo + tid; g < g_hi; g += kTileThreads) {
                const uint32_t f0 = 4 * g;
                uint32_t l = f0 / (uint32_t)K;
                uint32_t c = f0 - l * (uint32_t)K;
                float f[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t lab = (l < pl) ? buf8[base + l] : 0xFFFFFFFFu;
                    f[k] = (lab == c) ? 1.0f : 0.0f;
                    if (++c == (uint32_t)K) {
                        c = 0;
                        l++;
                    }
                }
                if (g < full) {
                    st_cs(reinterpret_cast<float4*>(dst) + g, make_float4(f[0], f[1], f[2], f[3]));
                } else {
                    for (uint32_t b = f0; b < nfl; b++) dst[b] = f[b - f0];
                }
            }
        }
    }
}

// x^(8*n) mod P via the x^(2^k) table
__device__ inline uint32_t xpow8(const CrcTables* tab, uint64_t n) {
    uint32_t p = 0x80000000u;
    int k = 3;
    while (n) {
        if (n & 1) p = multmodp(__ldg(&tab->x2n[k & 63]), p);
        n >>= 1;
        k++;
    }
    return p;
}

// Fold the per-tile partials of one record into its CRC-32C (one warp per record, all lanes return it).
__device__ inline uint32_t fold_record_crc(const uint32_t* tc, uint32_t nt, uint64_t d0, uint64_t len,
                                           const uint8_t* base, const CrcTables* tab) {
    const int lane = threadIdx.x & 31;
    if (len < 4) {  // the init XOR does not fit in the message: do it bytewise
        uint32_t s = 0xFFFFFFFFu;
        for (uint64_t i = 0; i < len; i++) s = (s >> 8) ^ __ldg(&tab->t4[3][(s ^ base[d0 + i]) & 0xff]);
        return ~s;
    }
    const uint32_t q = (nt + 31) / 32;
    const uint32_t b = lane * q, e = (b + q < nt) ? b + q : nt;
    const uint32_t xt = tab->xtile;
    uint32_t s = 0;
    for (uint32_t t = b; t < e; t++) s = multmodp(xt, s) ^ tc[t];
    // lane chunks: advance lane's partial past the tiles of the following lanes
    uint32_t acc = 0;
    const uint32_t xq = xpow8(tab, (uint64_t)q * kTile);
    for (int l = 0; l < 32; l++) {
        const uint32_t sl = __shfl_sync(0xffffffffu, s, l);
        const uint32_t bl = l * q;
        if (bl >= nt) break;
        const uint32_t el = (bl + q < nt) ? bl + q : nt;
        // Horner across lanes: previous accumulation moves forward by this lane's tile count
        acc = ((el - bl) == q ? multmodp(xq, acc) : multmodp(xpow8(tab, (uint64_t)(el - bl) * kTile), acc)) ^ sl;
    }
    // acc sits at the end of the last tile; un-advance by the zero padding after the record end
    const uint64_t A = d0 & ~15ull;
    const uint64_t pad = A + (uint64_t)nt * kTile - (d0 + len);
    acc = multmodp(__ldg(&tab->xinv16[pad >> 4]), acc);
    acc = multmodp(__ldg(&tab->xinvb[pad & 15]), acc);
    return ~acc;
}

__device__ __forceinline__ uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

struct FinalArgs {
    const uint8_t* shard;
    uint64_t nbytes;
    const uint64_t* rec_off;
    const uint64_t* rec_len;
    const b2_example_index* index;
    b2_parse_sink sink;
    const CrcTables* tab;
    const uint32_t* tilecrc;
    uint32_t tiles_x;
    int n;
    int32_t* status;    // parse: per-record status
    uint32_t* crc_out;  // b2_crc32c: raw CRCs
};

__global__ void __launch_bounds__(256) parse_final_kernel(const FinalArgs a) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    const uint64_t d0 = a.rec_off[r], len = a.rec_len[r];
    const uint64_t A = d0 & ~15ull;
    const uint32_t nt = len ? (uint32_t)((d0 + len - A + kTile - 1) / kTile) : 0;
    uint32_t crc = 0;
    if (a.sink.verify_crc || a.crc_out) crc = fold_record_crc(a.tilecrc + (size_t)r * a.tiles_x, nt, d0, len, a.shard, a.tab);
    if (lane != 0) return;
    if (a.crc_out) {
        a.crc_out[r] = crc;
        return;
    }
    int32_t st = 0;
    if (a.sink.verify_crc) {
        uint32_t stored = 0;
        if (d0 + len + 4 <= a.nbytes)
            for (int j = 0; j < 4; j++) stored |= (uint32_t)a.shard[d0 + len + j] << (8 * j);
        else
            stored = ~mask_crc(crc);
        if (stored != mask_crc(crc)) st = 1;
    }
    if (st == 0 && a.index && a.sink.mode != B2_SINK_NONE) {
        const b2_example_index ix = a.index[r];
        if (ix.status != 0) st = 2;
        else if (a.sink.mode == B2_SINK_RAW) {
            if (ix.img_len > a.sink.img_stride || ix.tgt_len > a.sink.tgt_stride) st = 3;
        } else {
            if (ix.img_kind != 1 || ix.tgt_kind != 1) st = 2;
            else if (ix.img_len * 4 > a.sink.img_stride || ix.tgt_len * (uint64_t)a.sink.num_classes * 4 > a.sink.tgt_stride ||
                     ix.img_len >= (1ull << 31) || ix.tgt_len * (uint64_t)a.sink.num_classes >= (1ull << 31)) st = 3;
        }
    }
    a.status[r] = st;
}

// ---------------------------------------------------------------------------------------------- scan
// Frames are a linked list (each length tells where the next header is).  Fast path: if the first record's
// stride divides the shard, every thread checks "its" header at i*stride; when all lengths agree the
// sequential walk would visit exactly those offsets (induction), so the result is identical.  Otherwise
// thread 0 walks the chain.
__device__ inline uint64_t rd_u64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}
__device__ inline uint32_t rd_u32(const uint8_t* p) {
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i);
    return v;
}
__device__ inline bool header_ok(const uint8_t* p, const CrcTables* tab) {
    uint32_t s = 0xFFFFFFFFu;
    for (int i = 0; i < 8; i++) s = (s >> 8) ^ __ldg(&tab->t4[3][(s ^ p[i]) & 0xff]);
    return mask_crc(~s) == rd_u32(p + 8);
}

__global__ void __launch_bounds__(1024)
scan_kernel(const uint8_t* __restrict__ shard, uint64_t nbytes, uint64_t cap, uint64_t* __restrict__ offs,
            uint64_t* __restrict__ lens, int64_t* __restrict__ result, const CrcTables* __restrict__ tab) {
    __shared__ int bad;
    __shared__ uint64_t s_stride;
    if (threadIdx.x == 0) {
        bad = 0;
        s_stride = 0;
        if (nbytes >= 16) {
            const uint64_t l0 = rd_u64(shard);
            if (l0 <= nbytes - 16 && (nbytes % (l0 + 16)) == 0) s_stride = l0 + 16;
        }
    }
    __syncthreads();
    const uint64_t stride = s_stride;
    if (nbytes == 0) {
        if (threadIdx.x == 0) { result[0] = 0; result[1] = 0; }
        return;
    }
    if (stride) {
        const uint64_t n = nbytes / stride;
        if (n <= cap) {
            for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint8_t* h = shard + i * stride;
                if (rd_u64(h) != stride - 16 || !header_ok(h, tab)) bad = 1;
            }
            __syncthreads();
            if (!bad) {
                for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
                    offs[i] = i * stride + 12;
                    lens[i] = stride - 16;
                }
                if (threadIdx.x == 0) { result[0] = (int64_t)n; result[1] = 0; }
                return;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint64_t pos = 0, n = 0;
    int64_t st = 0;
    while (pos < nbytes) {
        if (nbytes - pos < 12) { st = 1; break; }
        const uint64_t l = rd_u64(shard + pos);
        if (!header_ok(shard + pos, tab)) { st = 1; break; }
        if (l > nbytes - pos - 12 || nbytes - pos - 12 - l < 4) { st = 1; break; }
        if (n >= cap) { st = 2; break; }
        offs[n] = pos + 12;
        lens[n] = l;
        n++;
        pos += 16 + l;
    }
    result[0] = (int64_t)n;
    result[1] = st;
}

// ---------------------------------------------------------------------------------------------- index
// One thread per record walks the protobuf structure.  Everything is bounds-checked
// against the record; a malformed message sets status 1 instead of faulting.
struct Cursor {
    const uint8_t* b;
    uint64_t p, end;
    bool ok;
};
__device__ inline uint64_t rd_varint(Cursor& c) {
    uint64_t v = 0;
    for (int s = 0; s < 70; s += 7) {
        if (c.p >= c.end) { c.ok = false; return 0; }
        const uint8_t x = c.b[c.p++];
        v |= (uint64_t)(x & 0x7F) << s;
        if (!(x & 0x80)) return v;
    }
    c.ok = false;
    return 0;
}
// reads a tag and, for LEN fields, the sub-range; skips other wire types. returns field number (0 on end/error)
__device__ inline uint32_t next_field(Cursor& c, int& wt, uint64_t& v0, uint64_t& v1) {
    if (!c.ok || c.p >= c.end) return 0;
    const uint64_t tag = rd_varint(c);
    if (!c.ok) return 0;
    wt = (int)(tag & 7);
    const uint32_t f = (uint32_t)(tag >> 3);
    if (wt == 0) {
        v0 = rd_varint(c);
    } else if (wt == 1) {
        v0 = c.p; v1 = c.p + 8; c.p += 8;
    } else if (wt == 5) {
        v0 = c.p; v1 = c.p + 4; c.p += 4;
    } else if (wt == 2) {
        const uint64_t n = rd_varint(c);
        if (!c.ok || n > c.end - c.p) { c.ok = false; return 0; }
        v0 = c.p; v1 = c.p + n; c.p += n;
    } else {
        c.ok = false;
        return 0;
    }
    if (c.p > c.end) { c.ok = false; return 0; }
    if (f == 0) { c.ok = false; return 0; }
    return f;
}
__device__ inline bool key_is(const uint8_t* b, uint64_t s, uint64_t e, const char* lit, int n) {
    if (e - s != (uint64_t)n) return false;
    for (int i = 0; i < n; i++)
        if (b[s + i] != (uint8_t)lit[i]) return false;
    return true;
}

__global__ void __launch_bounds__(128)
index_kernel(const uint8_t* __restrict__ shard, const uint64_t* __restrict__ rec_off,
             const uint64_t* __restrict__ rec_len, int n, b2_example_index* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    b2_example_index ix;
    memset(&ix, 0, sizeof(ix));
    int have[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // img, h, w, c, tgt, th, tw, id : 1 ok, -1 wrong type/count
    int64_t dims[5] = {0, 0, 0, 0, 0};
    Cursor ex{shard, rec_off[r], rec_off[r] + rec_len[r], true};
    int wt; uint64_t a, b;
    while (uint32_t f = next_field(ex, wt, a, b)) {
        if (f != 1 || wt != 2) continue;  // Example.features
        Cursor fs{shard, a, b, true};
        int wt2; uint64_t a2, b2v;
        while (uint32_t f2 = next_field(fs, wt2, a2, b2v)) {
            if (f2 != 1 || wt2 != 2) continue;  // Features.feature map entry
            Cursor en{shard, a2, b2v, true};
            int wt3; uint64_t a3, b3;
            uint64_t ks = 0, ke = 0, vs = 0, ve = 0;
            bool hk = false, hv = false;
            while (uint32_t f3 = next_field(en, wt3, a3, b3)) {
                if (f3 == 1 && wt3 == 2) { ks = a3; ke = b3; hk = true; }
                else if (f3 == 2 && wt3 == 2) { vs = a3; ve = b3; hv = true; }
            }
            if (!en.ok) { ex.ok = false; break; }
            if (!hk) continue;
            int which = -1;
            if (key_is(shard, ks, ke, "image/image_data", 16)) which = 0;
            else if (key_is(shard, ks, ke, "image/height", 12)) which = 1;
            else if (key_is(shard, ks, ke, "image/width", 11)) which = 2;
            else if (key_is(shard, ks, ke, "image/channels", 14)) which = 3;
            else if (key_is(shard, ks, ke, "target/target_data", 18)) which = 4;
            else if (key_is(shard, ks, ke, "target/height", 13)) which = 5;
            else if (key_is(shard, ks, ke, "target/width", 12)) which = 6;
            else if (key_is(shard, ks, ke, "identifier", 10)) which = 7;
            if (which < 0) continue;
            // Feature oneof: last member present wins
            int kind = 0; uint64_t ls = 0, le = 0;
            if (hv) {
                Cursor fe{shard, vs, ve, true};
                int wt4; uint64_t a4, b4;
                while (uint32_t f4 = next_field(fe, wt4, a4, b4)) {
                    if (wt4 == 2 && f4 >= 1 && f4 <= 3) { kind = (int)f4; ls = a4; le = b4; }
                }
                if (!fe.ok) { ex.ok = false; break; }
            }
            // the list message: field 1 repeated
            uint64_t ps = 0, pe = 0; int count = 0; int64_t ival = 0; bool packed_ok = true;
            if (kind) {
                Cursor li{shard, ls, le, true};
                int wt5; uint64_t a5, b5;
                while (uint32_t f5 = next_field(li, wt5, a5, b5)) {
                    if (f5 != 1) continue;
                    if (kind == 1) {
                        if (wt5 == 2) { ps = a5; pe = b5; count++; }
                    } else if (kind == 2) {
                        if (wt5 == 2) { ps = a5; pe = b5; count++; }      // packed floats (one chunk expected)
                        else if (wt5 == 5) { packed_ok = false; }          // unpacked fixed32: not supported on device
                    } else {
                        if (wt5 == 2) {
                            Cursor pk{shard, a5, b5, true};
                            while (pk.ok && pk.p < pk.end) { ival = (int64_t)rd_varint(pk); count++; }
                            if (!pk.ok) li.ok = false;
                        } else if (wt5 == 0) { ival = (int64_t)a5; count++; }
                    }
                }
                if (!li.ok) { ex.ok = false; break; }
            }
            if (which == 0 || which == 4) {
                int okk = 0;
                if (kind == 1 && count == 1) okk = 1;
                else if (kind == 2 && packed_ok && count <= 1 && ((pe - ps) & 3) == 0) okk = 2;
                if (which == 0) { ix.img_off = ps; ix.img_len = pe - ps; ix.img_kind = okk; have[0] = okk ? 1 : -1; }
                else { ix.tgt_off = ps; ix.tgt_len = pe - ps; ix.tgt_kind = okk; have[4] = okk ? 1 : -1; }
            } else if (which == 7) {
                if (kind == 1 && count == 1) { ix.id_off = ps; ix.id_len = pe - ps; have[7] = 1; } else have[7] = -1;
            } else {
                const int d = which < 4 ? which - 1 : which - 2;  // h,w,c,th,tw -> 0..4
                if (kind == 3 && count == 1) { dims[d] = ival; have[which] = 1; } else have[which] = -1;
            }
        }
        if (!fs.ok) ex.ok = false;
        if (!ex.ok) break;
    }
    ix.height = (int32_t)dims[0]; ix.width = (int32_t)dims[1]; ix.channels = (int32_t)dims[2];
    ix.tgt_height = (int32_t)dims[3]; ix.tgt_width = (int32_t)dims[4];
    int st = ex.ok ? 0 : 1;
    if (st == 0)
        for (int k = 0; k < 8; k++)
            if (have[k] != 1) st = 2;
    ix.status = st;
    out[r] = ix;
}

// ---------------------------------------------------------------------------------------------- build
struct BuildArgs {
    const b2_build_desc* descs;
    const uint8_t* scaffold;
    uint8_t* out;
    const CrcTables* tab;
    uint32_t* tilecrc;
    uint32_t tiles_x;
    int n;
};

__device__ __forceinline__ uint32_t elem_as_f32_bits(const void* src, int dtype, uint64_t i, uint64_t count) {
    if (i >= count) return 0;
    float f;
    switch (dtype) {
        case B2_U8: f = (float)static_cast<const uint8_t*>(src)[i]; break;
        case B2_I8: f = (float)static_cast<const int8_t*>(src)[i]; break;
        case B2_U16: f = (float)static_cast<const uint16_t*>(src)[i]; break;
        case B2_I16: f = (float)static_cast<const int16_t*>(src)[i]; break;
        case B2_U32: f = (float)static_cast<const uint32_t*>(src)[i]; break;
        case B2_I32: f = (float)static_cast<const int32_t*>(src)[i]; break;
        case B2_F32: f = static_cast<const float*>(src)[i]; break;
        default: f = (float)static_cast<const double*>(src)[i]; break;
    }
    return __float_as_uint(f);
}

// byte j of a payload
__device__ __forceinline__ uint32_t payload_byte(const void* src, int kind, int dtype, uint64_t j, uint64_t count) {
    if (kind == 1) return static_cast<const uint8_t*>(src)[j];
    return (elem_as_f32_bits(src, dtype, j >> 2, count) >> (8 * (j & 3))) & 0xFFu;
}
// four consecutive payload bytes j..j+3 (all inside the payload)
__device__ __forceinline__ uint32_t payload_word(const void* src, int kind, int dtype, uint64_t j, uint64_t count) {
    if (kind == 1) {
        const uint8_t* p = static_cast<const uint8_t*>(src) + j;
        if ((reinterpret_cast<uintptr_t>(p) & 3) == 0) return *reinterpret_cast<const uint32_t*>(p);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(3));
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
        // the second aligned word may lie past the payload end only if j+4 == count and p unaligned: then its
        // needed bytes are still inside, so the aligned read stays within the allocation's last word
        return __funnelshift_r(q[0], q[1], sh);
    }
    const uint32_t b0 = elem_as_f32_bits(src, dtype, j >> 2, count);
    if ((j & 3) == 0) return b0;
    const uint32_t b1 = elem_as_f32_bits(src, dtype, (j >> 2) + 1, count);
    return __funnelshift_r(b0, b1, (uint32_t)(j & 3) * 8);
}

__global__ void __launch_bounds__(kTileThreads) build_kernel(const BuildArgs a) {
    __shared__ __align__(16) uint4 buf4[kTile / 16];
    __shared__ CrcSmem cs;
    __shared__ uint32_t red[kTileThreads / 32];
    __shared__ uint32_t hdr_crc;
    const int r = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    const b2_build_desc d = a.descs[r];
    const uint64_t rs = d.out_off, L = d.example_len;
    const uint64_t d0 = rs + 12, d1 = d0 + L, re = d1;  // footer (4 bytes at d1) is written by the final kernel
    const uint64_t A = rs & ~15ull;
    const uint64_t ts = A + (uint64_t)tile * kTile, te = ts + kTile;
    if (ts >= re) return;
    load_crc_tables(&cs, a.tab);
    if (threadIdx.x == 0) {
        uint32_t s = 0xFFFFFFFFu;
        for (int i = 0; i < 8; i++) s = (s >> 8) ^ __ldg(&a.tab->t4[3][(s ^ (uint32_t)((L >> (8 * i)) & 0xFF)) & 0xff]);
        hdr_crc = mask_crc(~s);
    }
    __syncthreads();
    const uint64_t ib = d.kind == 1 ? d.img_count : d.img_count * 4, tb = d.kind == 1 ? d.tgt_count : d.tgt_count * 4;
    // Example-space segment boundaries
    const uint64_t e1 = d.piece_len[0], e2 = e1 + ib, e3 = e2 + d.piece_len[1], e4 = e3 + tb;
    const uint8_t* sc = a.scaffold + d.scaffold_off;
    uint32_t* buf32 = reinterpret_cast<uint32_t*>(buf4);
    for (int w = threadIdx.x; w < kTile / 4; w += blockDim.x) {
        const uint64_t p = ts + 4ull * w;  // absolute position of this word
        uint32_t word = 0;
        if (p + 4 > rs && p < re) {
            const int64_t e = (int64_t)p - (int64_t)d0;  // Example-space offset of byte 0
            if (e >= (int64_t)e1 && (uint64_t)e + 4 <= e2) {
                word = payload_word(d.img_src, d.kind, d.src_dtype, (uint64_t)e - e1, d.img_count);
            } else if (e >= (int64_t)e3 && (uint64_t)e + 4 <= e4) {
                word = payload_word(d.tgt_src, d.kind, d.tgt_dtype, (uint64_t)e - e3, d.tgt_count);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint64_t q = p + j;
                    if (q < rs || q >= re) continue;
                    uint32_t bv;
                    const uint64_t rel = q - rs;
                    if (rel < 8) bv = (uint32_t)((L >> (8 * rel)) & 0xFF);
                    else if (rel < 12) bv = (hdr_crc >> (8 * (rel - 8))) & 0xFF;
                    else {
                        const uint64_t x = rel - 12;
                        if (x < e1) bv = sc[x];
                        else if (x < e2) bv = payload_byte(d.img_src, d.kind, d.src_dtype, x - e1, d.img_count);
                        else if (x < e3) bv = sc[e1 + (x - e2)];
                        else if (x < e4) bv = payload_byte(d.tgt_src, d.kind, d.tgt_dtype, x - e3, d.tgt_count);
                        else bv = sc[e1 + d.piece_len[1] + (x - e4)];
                    }
                    word |= bv << (8 * j);
                }
            }
        }
        buf32[w] = word;
    }
    __syncthreads();
    // write out the bytes of [max(ts,rs), min(te,re))
    const uint8_t* buf8 = reinterpret_cast<const uint8_t*>(buf4);
    for (int k = threadIdx.x; k < kTile / 16; k += blockDim.x) {
        const uint64_t p = ts + 16ull * k;
        if (p >= rs && p + 16 <= re) {
            st_cs(reinterpret_cast<uint4*>(a.out + p), buf4[k]);
        } else if (p + 16 > rs && p < re) {
            for (int j = 0; j < 16; j++)
                if (p + j >= rs && p + j < re) a.out[p + j] = buf8[16 * k + j];
        }
    }
    // CRC partial over the data range.  Tile grid here is anchored at A = rs & ~15 (not d0 & ~15); the final
    // kernel is told so through the same anchor.
    const uint32_t c = tile_crc(buf4, &cs, a.tab, ts, d0, d1, /*first_tile=*/ts <= d0 && d0 < te, red);
    if (threadIdx.x == 0) a.tilecrc[(size_t)r * a.tiles_x + tile] = c;
}

__global__ void __launch_bounds__(256) build_final_kernel(const BuildArgs a) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= a.n) return;
    const b2_build_desc d = a.descs[r];
    const uint64_t rs = d.out_off, L = d.example_len, d0 = rs + 12;
    const uint64_t A = rs & ~15ull;
    // tiles containing data: first data tile index t0, count nt — fold only those, anchored at A + t0*kTile
    const uint32_t t0 = (uint32_t)((d0 - A) / kTile);
    uint32_t crc;
    if (L < 4) {
        uint32_t s = 0xFFFFFFFFu;
        for (uint64_t i = 0; i < L; i++) s = (s >> 8) ^ __ldg(&a.tab->t4[3][(s ^ a.out[d0 + i]) & 0xff]);
        crc = ~s;
    } else {
        const uint32_t nt = (uint32_t)((d0 + L - A + kTile - 1) / kTile) - t0;
        // fold_record_crc derives its anchor from d0 & ~15; emulate by folding manually with our anchor
        const uint32_t* tc = a.tilecrc + (size_t)r * a.tiles_x + t0;
        const uint32_t q = (nt + 31) / 32;
        const uint32_t b = lane * q, e = (b + q < nt) ? b + q : nt;
        uint32_t s = 0;
        for (uint32_t t = b; t < e; t++) s = multmodp(a.tab->xtile, s) ^ tc[t];
        uint32_t acc = 0;
        const uint32_t xq = xpow8(a.tab, (uint64_t)q * kTile);
        for (int l = 0; l < 32; l++) {
            const uint32_t sl = __shfl_sync(0xffffffffu, s, l);
            const uint32_t bl = l * q;
            if (bl >= nt) break;
            const uint32_t el = (bl + q < nt) ? bl + q : nt;
            acc = ((el - bl) == q ? multmodp(xq, acc) : multmodp(xpow8(a.tab, (uint64_t)(el - bl) * kTile), acc)) ^ sl;
        }
        const uint64_t pad = A + (uint64_t)(t0 + nt) * kTile - (d0 + L);
        acc = multmodp(__ldg(&a.tab->xinv16[pad >> 4]), acc);
        acc = multmodp(__ldg(&a.tab->xinvb[pad & 15]), acc);
        crc = ~acc;
    }
    if (lane == 0) {
        const uint32_t m = mask_crc(crc);
        for (int j = 0; j < 4; j++) a.out[d0 + L + j] = (uint8_t)(m >> (8 * j));
    }
}

}  // namespace b2

using namespace b2;

static int launch_parse(b2_ctx* ctx, const ParseArgs& pa, int n, uint32_t tiles_x, cudaStream_t s) {
    dim3 grid(tiles_x, n);
    switch (pa.sink.mode) {
        case B2_SINK_NONE: parse_kernel<B2_SINK_NONE><<<grid, kTileThreads, 0, s>>>(pa); break;
        case B2_SINK_RAW: parse_kernel<B2_SINK_RAW><<<grid, kTileThreads, 0, s>>>(pa); break;
        default: {
            const int C = pa.sink.channels, K = pa.sink.num_classes;
            size_t dyn = (C <= kLutMaxC ? (size_t)C * 256 * sizeof(float) : 0) +
                         (K <= kHotMaxK ? (size_t)hot_labels_per_iter(K) * K * (kTileThreads / 32) * sizeof(float) : 0);
            static bool attr_set[64] = {false};
            if (!attr_set[ctx->device & 63]) {
                B2_CUDA(cudaFuncSetAttribute(parse_kernel<B2_SINK_NORM_ONEHOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
                attr_set[ctx->device & 63] = true;
            }
            parse_kernel<B2_SINK_NORM_ONEHOT><<<grid, kTileThreads, dyn, s>>>(pa);
        } break;
    }
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b2_tfrecord_parse(b2_ctx* ctx, const uint8_t* shard, uint64_t nbytes, const uint64_t* rec_off,
                                 const uint64_t* rec_len, const b2_example_index* index, int n,
                                 uint64_t max_record_len, const b2_parse_sink* sink, int32_t* status,
                                 b2_stream stream) {
    B2_REQUIRE(ctx && shard && rec_off && rec_len && sink && status, "b2_tfrecord_parse: NULL argument");
    B2_REQUIRE(n >= 0 && n <= 65535, "b2_tfrecord_parse: n must be in [0,65535] per call");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(shard) & 15) == 0, "b2_tfrecord_parse: shard must be 16-byte aligned");
    B2_REQUIRE(sink->mode == B2_SINK_NONE || index, "b2_tfrecord_parse: index required for a payload sink");
    if (sink->mode == B2_SINK_NORM_ONEHOT) {
        B2_REQUIRE(sink->mean && sink->std && sink->channels >= 1 && sink->channels <= 64 && sink->num_classes >= 1,
                   "b2_tfrecord_parse: NORM_ONEHOT needs mean/std, 1..64 channels and num_classes >= 1");
        B2_REQUIRE((reinterpret_cast<uintptr_t>(sink->img_out) & 15) == 0 && (sink->img_stride & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(sink->tgt_out) & 15) == 0 && (sink->tgt_stride & 15) == 0,
                   "b2_tfrecord_parse: NORM_ONEHOT outputs and strides must be 16-byte aligned");
    }
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint32_t tiles_x = (uint32_t)((max_record_len + 15 + kTile - 1) / kTile) + 0;
    B2_REQUIRE(tiles_x >= 1 || max_record_len == 0, "b2_tfrecord_parse: bad max_record_len");
    const uint32_t tx = tiles_x ? tiles_x : 1;
    if (int e = ws_reserve(ctx, (size_t)n * tx * sizeof(uint32_t), s)) return e;
    ParseArgs pa{shard, nbytes, rec_off, rec_len, index, *sink, ctx->crc_dev, static_cast<uint32_t*>(ctx->ws), tx};
    if (int e = launch_parse(ctx, pa, n, tx, s)) return e;
    FinalArgs fa{shard, nbytes, rec_off, rec_len, index, *sink, ctx->crc_dev, static_cast<uint32_t*>(ctx->ws), tx, n, status, nullptr};
    parse_final_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(fa);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b2_crc32c(b2_ctx* ctx, const uint8_t* data, const uint64_t* offsets, const uint64_t* lens, int n,
                         uint64_t max_len, uint32_t* crc_out, b2_stream stream) {
    B2_REQUIRE(ctx && data && offsets && lens && crc_out, "b2_crc32c: NULL argument");
    B2_REQUIRE(n >= 0 && n <= 65535, "b2_crc32c: n must be in [0,65535] per call");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(data) & 15) == 0, "b2_crc32c: data must be 16-byte aligned");
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint32_t tx = (uint32_t)((max_len + 15 + kTile - 1) / kTile);
    if (!tx) tx = 1;
    if (int e = ws_reserve(ctx, (size_t)n * tx * sizeof(uint32_t), s)) return e;
    b2_parse_sink sink;
    memset(&sink, 0, sizeof(sink));
    sink.mode = B2_SINK_NONE;
    sink.verify_crc = 1;
    ParseArgs pa{data, ~0ull, offsets, lens, nullptr, sink, ctx->crc_dev, static_cast<uint32_t*>(ctx->ws), tx};
    if (int e = launch_parse(ctx, pa, n, tx, s)) return e;
    FinalArgs fa{data, ~0ull, offsets, lens, nullptr, sink, ctx->crc_dev, static_cast<uint32_t*>(ctx->ws), tx, n, nullptr, crc_out};
    parse_final_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(fa);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b2_tfrecord_scan(b2_ctx* ctx, const uint8_t* shard, uint64_t nbytes, uint64_t max_records,
                                uint64_t* rec_off, uint64_t* rec_len, int64_t* result, b2_stream stream) {
    B2_REQUIRE(ctx && (shard || nbytes == 0) && rec_off && rec_len && result, "b2_tfrecord_scan: NULL argument");
    DeviceGuard g(ctx->device);
    scan_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(shard, nbytes, max_records, rec_off, rec_len, result, ctx->crc_dev);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b2_tfrecord_index(b2_ctx* ctx, const uint8_t* shard, const uint64_t* rec_off, const uint64_t* rec_len,
                                 int n, b2_example_index* out, b2_stream stream) {
    B2_REQUIRE(ctx && shard && rec_off && rec_len && out, "b2_tfrecord_index: NULL argument");
    B2_REQUIRE(n >= 0, "b2_tfrecord_index: n < 0");
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    index_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(shard, rec_off, rec_len, n, out);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int b2_tfrecord_build(b2_ctx* ctx, const b2_build_desc* descs, int n, uint64_t max_record_bytes,
                                 const uint8_t* scaffold, uint8_t* out, b2_stream stream) {
    B2_REQUIRE(ctx && descs && scaffold && out, "b2_tfrecord_build: NULL argument");
    B2_REQUIRE(n >= 0 && n <= 65535, "b2_tfrecord_build: n must be in [0,65535] per call");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "b2_tfrecord_build: out must be 16-byte aligned");
    if (n == 0) return 0;
    DeviceGuard g(ctx->device);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint32_t tx = (uint32_t)((max_record_bytes + 15 + kTile - 1) / kTile);
    if (!tx) tx = 1;
    if (int e = ws_reserve(ctx, (size_t)n * tx * sizeof(uint32_t), s)) return e;
    BuildArgs ba{descs, scaffold, out, ctx->crc_dev, static_cast<uint32_t*>(ctx->ws), tx, n};
    build_kernel<<<dim3(tx, n), kTileThreads, 0, s>>>(ba);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    build_final_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(ba);
    ctx->launches++;
    B2_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- host: Example layout
namespace {
void put_varint(std::vector<uint8_t>& v, uint64_t x) {
    while (x >= 0x80) {
        v.push_back((uint8_t)(x | 0x80));
        x >>= 7;
    }
    v.push_back((uint8_t)x);
}
size_t varint_len(uint64_t x) {
    size_t n = 1;
    while (x >= 0x80) { x >>= 7; n++; }
    return n;
}
// entry header for a payload feature: everything up to (not including) the payload bytes
// entry = 0a <elen> [ 0a <klen> key  12 <flen> [ <ftag> <llen> [ 0a <plen> payload ] ] ]
void payload_entry_head(std::vector<uint8_t>& v, const char* key, int kind, uint64_t plen) {
    const size_t klen = strlen(key);
    const uint64_t list_len = plen ? 1 + varint_len(plen) + plen : (kind == 1 ? 2 : 0);
    const uint64_t feat_len = 1 + varint_len(list_len) + list_len;
    const uint64_t entry_len = 1 + varint_len(klen) + klen + 1 + varint_len(feat_len) + feat_len;
    v.push_back(0x0a); put_varint(v, entry_len);
    v.push_back(0x0a); put_varint(v, klen); v.insert(v.end(), key, key + klen);
    v.push_back(0x12); put_varint(v, feat_len);
    v.push_back(kind == 1 ? 0x0a : 0x12); put_varint(v, list_len);
    if (plen || kind == 1) { v.push_back(0x0a); put_varint(v, plen); }
}
uint64_t payload_entry_total(const char* key, int kind, uint64_t plen) {
    std::vector<uint8_t> t;
    payload_entry_head(t, key, kind, plen);
    return t.size() + plen;
}
void int_entry(std::vector<uint8_t>& v, const char* key, int64_t val) {
    const size_t klen = strlen(key);
    const size_t vl = varint_len((uint64_t)val);
    const uint64_t list_len = 1 + 1 + vl, feat_len = 1 + 1 + list_len;
    const uint64_t entry_len = 1 + varint_len(klen) + klen + 1 + 1 + feat_len;
    v.push_back(0x0a); put_varint(v, entry_len);
    v.push_back(0x0a); put_varint(v, klen); v.insert(v.end(), key, key + klen);
    v.push_back(0x12); put_varint(v, feat_len);
    v.push_back(0x1a); put_varint(v, list_len);
    v.push_back(0x0a); put_varint(v, vl); put_varint(v, (uint64_t)val);
}
void bytes_entry(std::vector<uint8_t>& v, const char* key, const uint8_t* data, uint64_t n) {
    payload_entry_head(v, key, 1, n);
    v.insert(v.end(), data, data + n);
}
}  // namespace

extern "C" int b2_example_layout(int kind, uint64_t img_bytes, uint64_t tgt_bytes, int64_t img_h, int64_t img_w,
                                 int64_t img_c, int64_t tgt_h, int64_t tgt_w, const uint8_t* identifier,
                                 uint64_t identifier_len, uint8_t* scaffold, uint64_t cap, uint32_t piece_len[3],
                                 uint64_t* example_len) {
    B2_REQUIRE(kind == 1 || kind == 2, "b2_example_layout: kind must be 1 (BytesList) or 2 (FloatList)");
    B2_REQUIRE(scaffold && piece_len && example_len && (identifier || identifier_len == 0), "b2_example_layout: NULL argument");
    B2_REQUIRE(kind == 1 || ((img_bytes | tgt_bytes) & 3) == 0, "b2_example_layout: FloatList payload must be a multiple of 4 bytes");
    // sorted key order: identifier, image/channels, image/height, image/image_data, image/width,
    //                   target/height, target/target_data, target/width
    std::vector<uint8_t> p0, p1, p2;
    bytes_entry(p0, "identifier", identifier, identifier_len);
    int_entry(p0, "image/channels", img_c);
    int_entry(p0, "image/height", img_h);
    payload_entry_head(p0, "image/image_data", kind, img_bytes);
    int_entry(p1, "image/width", img_w);
    int_entry(p1, "target/height", tgt_h);
    payload_entry_head(p1, "target/target_data", kind, tgt_bytes);
    int_entry(p2, "target/width", tgt_w);
    const uint64_t features_len = p0.size() + img_bytes + p1.size() + tgt_bytes + p2.size();
    std::vector<uint8_t> head;
    head.push_back(0x0a);
    put_varint(head, features_len);
    p0.insert(p0.begin(), head.begin(), head.end());
    B2_REQUIRE(p0.size() + p1.size() + p2.size() <= cap, "b2_example_layout: scaffold buffer too small");
    memcpy(scaffold, p0.data(), p0.size());
    memcpy(scaffold + p0.size(), p1.data(), p1.size());
    memcpy(scaffold + p0.size() + p1.size(), p2.data(), p2.size());
    piece_len[0] = (uint32_t)p0.size();
    piece_len[1] = (uint32_t)p1.size();
    piece_len[2] = (uint32_t)p2.size();
    *example_len = p0.size() + img_bytes + p1.size() + tgt_bytes + p2.size();
    (void)payload_entry_total;
    return 0;
}
