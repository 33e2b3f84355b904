/*
 * oracle/csrc/b2oracle.c — CPU restatement helpers.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (dl_image_segmentation_b200) never does.
 *
 * The reference (harry-gibson/dl_image_segmentation) delegates every routine here to
 * un-vendored third-party native code, so each function restates the PUBLISHED
 * algorithm and cites the reference call site whose behaviour it stands in for:
 *
 *   orc_crc32c / orc_masked_crc32c / orc_tfrecord_*   TensorFlow RecordWriter/Reader reached from
 *        tf.io.TFRecordWriter.write   (_img_to_tf_mp.py:119,141  _img_to_tf_threaded.py:182,203)
 *        tf.data.TFRecordDataset      (parse_tfrecords.ipynb cell 4)
 *        CRC-32C: RFC 3720 App. B.4, reflected poly 0x82F63B78; mask per TF crc32c.h.
 *   orc_lzw_decode   libtiff LZW reached through rasterio MemoryFile(...).read()
 *        (_img_to_tf_mp.py:45-48, _tfrecord_image_translation.py:320-326,369-381);
 *        TIFF 6.0 section 13: MSB-first codes, Clear=256, EOI=257, early change.
 *   orc_png_unfilter libpng row un-filter reached through tf.image.decode_png
 *        (_img_to_tf_threaded.py:59) / tf.io.decode_image (_tfrecord_image_translation.py:283,289);
 *        PNG spec section 9 (None/Sub/Up/Average/Paeth).
 *   orc_hdiff_undo   TIFF 6.0 section 14 predictor 2 (horizontal differencing).
 *
 * Parity status: the reference holds no golden vectors for any of these;
 * tests pin these routines against RFC vectors, zlib, libtiff (via cv2 / Pillow) and libpng.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>

#if defined(__SSE4_2__)
#include <nmmintrin.h>
#endif

/* ------------------------------------------------------------------ CRC-32C */

static uint32_t g_crc_tab[8][256];
static int g_crc_ready = 0;

static void crc_init(void) {
    if (g_crc_ready) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
        g_crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xff];
    g_crc_ready = 1;
}

/* plain table version: the arithmetic definition (always available) */
uint32_t orc_crc32c_sw(const uint8_t* p, size_t n) {
    crc_init();
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = (c >> 8) ^ g_crc_tab[0][(c ^ p[i]) & 0xff];
    return c ^ 0xFFFFFFFFu;
}

/* what TF actually executes on x86: the SSE4.2 crc32 instruction */
uint32_t orc_crc32c(const uint8_t* p, size_t n) {
#if defined(__SSE4_2__)
    uint64_t c = 0xFFFFFFFFu;
    while (n && ((uintptr_t)p & 7)) { c = _mm_crc32_u8((uint32_t)c, *p++); n--; }
    while (n >= 8) { uint64_t v; memcpy(&v, p, 8); c = _mm_crc32_u64(c, v); p += 8; n -= 8; }
    while (n) { c = _mm_crc32_u8((uint32_t)c, *p++); n--; }
    return (uint32_t)c ^ 0xFFFFFFFFu;
#else
    return orc_crc32c_sw(p, n);
#endif
}

uint32_t orc_mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }
uint32_t orc_masked_crc32c(const uint8_t* p, size_t n) { return orc_mask_crc(orc_crc32c(p, n)); }

/* Frame one record: u64le len | u32le maskedcrc(len) | data | u32le maskedcrc(data). out has n+16 bytes. */
void orc_tfrecord_frame(const uint8_t* data, uint64_t n, uint8_t* out) {
    uint8_t hdr[8];
    for (int i = 0; i < 8; i++) hdr[i] = (uint8_t)(n >> (8 * i));
    memcpy(out, hdr, 8);
    uint32_t c = orc_masked_crc32c(hdr, 8);
    for (int i = 0; i < 4; i++) out[8 + i] = (uint8_t)(c >> (8 * i));
    memcpy(out + 12, data, n);
    c = orc_masked_crc32c(data, n);
    for (int i = 0; i < 4; i++) out[12 + n + i] = (uint8_t)(c >> (8 * i));
}

/* Sequential scan as RecordReader does it.  Returns number of records, or
 * -(1+i) if record i is corrupt (bad length crc, truncated, or bad data crc when verify!=0). */
int64_t orc_tfrecord_scan(const uint8_t* buf, uint64_t nbytes, uint64_t* offs, uint64_t* lens,
                          int64_t cap, int verify) {
    uint64_t pos = 0; int64_t n = 0;
    while (pos < nbytes) {
        if (nbytes - pos < 12) return -(1 + n);
        uint64_t len = 0; uint32_t c = 0;
        for (int i = 0; i < 8; i++) len |= (uint64_t)buf[pos + i] << (8 * i);
        for (int i = 0; i < 4; i++) c |= (uint32_t)buf[pos + 8 + i] << (8 * i);
        if (orc_masked_crc32c(buf + pos, 8) != c) return -(1 + n);
        if (len > nbytes - pos - 12 || nbytes - pos - 12 - len < 4) return -(1 + n);
        if (verify) {
            uint32_t d = 0;
            for (int i = 0; i < 4; i++) d |= (uint32_t)buf[pos + 12 + len + i] << (8 * i);
            if (orc_masked_crc32c(buf + pos + 12, len) != d) return -(1 + n);
        }
        if (n < cap) { offs[n] = pos + 12; lens[n] = len; }
        n++; pos += 16 + len;
    }
    return n;
}

/* ------------------------------------------------------------------ TIFF LZW */

/* Classic table-driven decoder in the libtiff style (prefix chain, written back to front).
 * Returns bytes produced (== dst_len on a well-formed stream), or -1 on a corrupt stream. */
int64_t orc_lzw_decode(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len) {
    enum { CLEAR = 256, EOI = 257, FIRST = 258, MAXTAB = 4096 };
    static __thread uint16_t prefix[MAXTAB];
    static __thread uint8_t suffix[MAXTAB];
    static __thread uint8_t firstc[MAXTAB];
    static __thread uint16_t length[MAXTAB];
    for (int i = 0; i < 256; i++) { prefix[i] = 0xFFFF; suffix[i] = (uint8_t)i; firstc[i] = (uint8_t)i; length[i] = 1; }
    size_t bitpos = 0, total_bits = src_len * 8, out = 0;
    int nbits = 9, free_ent = FIRST, old = -1;
    while (out < dst_len) {
        if (bitpos + nbits > total_bits) break;            /* ran out of input: libtiff stops here */
        uint32_t code = 0;
        {   /* MSB-first extraction */
            size_t byte = bitpos >> 3; int sh = (int)(bitpos & 7);
            uint32_t w = 0;
            for (int k = 0; k < 4; k++) w = (w << 8) | (byte + k < src_len ? src[byte + k] : 0);
            code = (w << sh) >> (32 - nbits);
            bitpos += nbits;
        }
        if (code == EOI) break;
        if (code == CLEAR) { nbits = 9; free_ent = FIRST; old = -1; continue; }
        if (old < 0) {                                       /* first code after Clear is a literal */
            if (code > 255) return -1;
            dst[out++] = (uint8_t)code; old = (int)code; continue;
        }
        if ((int)code > free_ent || (free_ent >= MAXTAB && (int)code >= free_ent)) return -1;
        /* add the new entry: string(old) + firstchar(string(code)) ; KwKwK when code==free_ent */
        if (free_ent < MAXTAB) {
            prefix[free_ent] = (uint16_t)old;
            firstc[free_ent] = firstc[old];
            length[free_ent] = (uint16_t)(length[old] + 1);
            suffix[free_ent] = ((int)code < free_ent) ? firstc[code] : firstc[old];
            free_ent++;
            /* early change: widen when the NEXT free slot exceeds (1<<nbits)-2 */
            if (free_ent > (1 << nbits) - 2 && nbits < 12) nbits++;
        }
        {   /* emit string(code), back to front, truncated at dst_len like libtiff */
            size_t len = length[code];
            size_t room = dst_len - out;
            int c = (int)code;
            size_t skip = len > room ? len - room : 0;
            size_t w = len > room ? room : len;
            while (skip--) c = prefix[c];
            for (size_t i = w; i-- > 0;) { dst[out + i] = suffix[c]; c = prefix[c]; }
            out += w;
        }
        old = (int)code;
    }
    return (int64_t)out;
}

/* ------------------------------------------------------------------ predictors / filters */

/* TIFF predictor 2 undo, in place. rows x (row_samples) samples, stride spp, 8- or 16-bit words
 * (16-bit words already in host byte order). */
void orc_hdiff_undo(uint8_t* buf, size_t rows, size_t row_samples, int spp, int bytes_per_sample) {
    if (bytes_per_sample == 1) {
        for (size_t r = 0; r < rows; r++) {
            uint8_t* p = buf + r * row_samples;
            for (size_t i = (size_t)spp; i < row_samples; i++) p[i] = (uint8_t)(p[i] + p[i - spp]);
        }
    } else if (bytes_per_sample == 2) {
        for (size_t r = 0; r < rows; r++) {
            uint16_t* p = (uint16_t*)buf + r * row_samples;
            for (size_t i = (size_t)spp; i < row_samples; i++) p[i] = (uint16_t)(p[i] + p[i - spp]);
        }
    } else if (bytes_per_sample == 4) {
        for (size_t r = 0; r < rows; r++) {
            uint32_t* p = (uint32_t*)buf + r * row_samples;
            for (size_t i = (size_t)spp; i < row_samples; i++) p[i] = p[i] + p[i - spp];
        }
    }
}

/* PNG un-filter: src = h rows of (1 + rowbytes); dst = h*rowbytes. Returns 0, or -1 on bad filter type. */
int orc_png_unfilter(const uint8_t* src, uint8_t* dst, size_t h, size_t rowbytes, int bpp) {
    for (size_t y = 0; y < h; y++) {
        const uint8_t* s = src + y * (rowbytes + 1);
        int ft = s[0]; s++;
        uint8_t* d = dst + y * rowbytes;
        const uint8_t* up = y ? d - rowbytes : NULL;
        if (ft > 4) return -1;
        for (size_t x = 0; x < rowbytes; x++) {
            int a = x >= (size_t)bpp ? d[x - bpp] : 0;
            int b = up ? up[x] : 0;
            int c = (up && x >= (size_t)bpp) ? up[x - bpp] : 0;
            int v = s[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                } break;
            }
            d[x] = (uint8_t)v;
        }
    }
    return 0;
}

/* FloatList helper: widen u16 / u8 to little-endian float32 (what arr.flatten() -> FloatList does,
 * _tfrecord_image_translation.py:27,35). */
void orc_u16_to_f32(const uint16_t* s, float* d, size_t n) { for (size_t i = 0; i < n; i++) d[i] = (float)s[i]; }
void orc_u8_to_f32(const uint8_t* s, float* d, size_t n) { for (size_t i = 0; i < n; i++) d[i] = (float)s[i]; }
