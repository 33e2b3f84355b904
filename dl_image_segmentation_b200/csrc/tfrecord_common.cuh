// tfrecord_common.cuh — device helpers shared by the TFRecord kernels (tfrecord.cu: scan / index / build;
// parse.cu: the fused CRC + payload pass).
#pragma once
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

// a(x)*b(x) mod P, fully unrolled and branch-free (5 instructions per bit instead of the rolled loop's 11)
__device__ __forceinline__ uint32_t multmodp_fast(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll
    for (int k = 31; k >= 0; k--) {
        if (a & (1u << k)) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ kPoly : (b >> 1);
    }
    return p;
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
    s = multmodp_fast(__ldg(&tab->fix[i]), s);
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
__device__ __forceinline__ uint32_t smem_u32_unaligned(const uint32_t* buf32, uint32_t off) {
    const uint32_t w0 = buf32[off >> 2], w1 = buf32[(off >> 2) + 1];
    return __funnelshift_r(w0, w1, (off & 3) * 8);
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
__device__ __forceinline__ uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

// Per-thread running CRC state over consecutive tiles of one record.  A thread owns vectors i and i+256 of every tile; the
// distance from the end of one of its vectors to the start of its next one is always 4080 bytes, inside a tile and
// from one tile to the next, so the same "consume 4 bytes and skip 4080" table step chains them all and the
// expensive per-thread alignment (one GF(2)[x] multiplication) is paid once per run of tiles, not once per tile.
// On return the state sits at (tile end + 16 i): the caller un-advances by 16 i (xinv16[i]) when the run ends.
// ts = absolute start of the tile; [d0, d1) = the record's data; first = the tile holding d0 (the 0xFFFFFFFF init is
// XORed into the first four data bytes); interior = no byte of the tile lies outside [d0, d1).
__device__ __forceinline__ uint32_t crc_running_step(uint32_t s, const uint4* buf4, const CrcSmem* cs, uint64_t ts, uint64_t d0,
                                                     uint64_t d1, bool first, bool interior) {
    const int i = threadIdx.x;
    uint4 v0 = buf4[i], v1 = buf4[i + 256];
    if (!interior) {
        v0 = mask_vec(v0, ts + 16ull * i, d0, d1);
        v1 = mask_vec(v1, ts + 4096 + 16ull * i, d0, d1);
        if (first && ts + 16ull * i + 16 > d0 && ts + 16ull * i < d0 + 4) {   // vectors touching the first four data bytes
            uint32_t w[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t p = ts + 16ull * i + 4 * q + k;
                    if (p >= d0 && p < d0 + 4) w[q] ^= 0xFFu << (8 * k);
                }
            v0 = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    s = adv4(cs->t4, s ^ v0.x);
    s = adv4(cs->t4, s ^ v0.y);
    s = adv4(cs->t4, s ^ v0.z);
    s = adv4(cs->s, s ^ v0.w);
    s = adv4(cs->t4, s ^ v1.x);
    s = adv4(cs->t4, s ^ v1.y);
    s = adv4(cs->t4, s ^ v1.z);
    s = adv4(cs->s, s ^ v1.w);
    return s;
}

// x^(8*8192*j): advance a tile partial by j tiles
__device__ __forceinline__ uint32_t tile_power(const CrcTables* tab, uint32_t j) {
    return j < 2048 ? __ldg(&tab->tpow[j]) : xpow8(tab, (uint64_t)j * kTile);
}

// 16 bytes at shard + a (a is 16-aligned), zero beyond nbytes
__device__ __forceinline__ uint4 ld16_bounded(const uint8_t* shard, uint64_t a, uint64_t nbytes) {
    if (a + 16 <= nbytes) return ld_nc(reinterpret_cast<const uint4*>(shard + a));
    uint32_t w[4] = {0, 0, 0, 0};
    for (uint64_t p = a; p < nbytes; p++) w[(p - a) >> 2] |= (uint32_t)shard[p] << (8 * ((p - a) & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__host__ __device__ __forceinline__ uint32_t record_tiles(uint64_t d0, uint64_t len) {
    if (len == 0) return 1;                                  // an empty record still owns one (empty) tile
    return (uint32_t)((d0 + len - (d0 & ~15ull) + kTile - 1) / kTile);
}

// ---------------------------------------------------------------- shard table (b2_tfrecord_open / _parse_table)
// One caller-owned device buffer describing an opened shard; every section 16-byte aligned:
//   hdr int64[8]        [0] records  [1] scan status  [2] tiles  [3] longest record  [4] records with status != 0
//                       [5] chunk scheduler words of the fused pass
//   rec_off uint64[cap] | rec_len uint64[cap] | index b2_example_index[cap] | tile_start uint32[cap+1]
//   acc uint64[cap]                          (scratch of the fused pass: CRC partial | tiles done; zero between launches)
//   tile2rec uint32[cap_tiles]               (owner record of every 8 KiB tile)
struct TableView {
    int64_t* hdr;
    uint64_t* rec_off;
    uint64_t* rec_len;
    b2_example_index* index;
    uint32_t* tile_start;
    unsigned long long* acc;
    uint32_t* tile2rec;
    uint64_t cap, cap_tiles, bytes;
};
__host__ __device__ inline TableView table_view(uint8_t* t, uint64_t nbytes, uint64_t cap) {
    TableView v;
    auto up = [](uint64_t x) { return (x + 15) & ~15ull; };
    uint64_t o = 0;
    v.cap = cap;
    v.cap_tiles = nbytes / kTile + 2 * cap + 2;   // a record wastes at most two partial tiles
    v.hdr = reinterpret_cast<int64_t*>(t + o);             o += 64;
    v.rec_off = reinterpret_cast<uint64_t*>(t + o);        o += up(8 * cap);
    v.rec_len = reinterpret_cast<uint64_t*>(t + o);        o += up(8 * cap);
    v.index = reinterpret_cast<b2_example_index*>(t + o);  o += up(sizeof(b2_example_index) * cap);
    v.tile_start = reinterpret_cast<uint32_t*>(t + o);     o += up(4 * (cap + 1));
    v.acc = reinterpret_cast<unsigned long long*>(t + o);  o += up(8 * cap);
    v.tile2rec = reinterpret_cast<uint32_t*>(t + o);       o += up(4 * v.cap_tiles);
    v.bytes = o;
    return v;
}
__host__ __device__ inline uint64_t table_bytes(uint64_t nbytes, uint64_t cap) { return table_view(nullptr, nbytes, cap).bytes; }

}  // namespace b2
