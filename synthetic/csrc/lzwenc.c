/*
 * synthetic/csrc/lzwenc.c — TIFF-LZW *encoder* for writing synthetic chips (workload generator;
 * neither product nor oracle).  Mirrors what libtiff's encoder emits so the fixtures look like the
 * files GDAL writes for the reference (_descartes_img_chips.py:784 COMPRESS=LZW): MSB-first codes,
 * leading Clear, 9->12 bit "early change" widths, Clear when the table reaches 4094 entries, EOI.
 * Dictionary = first-child / next-sibling trie (a different structure from the decoders under test).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

typedef struct { uint64_t acc; int nacc; uint8_t* dst; size_t o, cap; int fail; } bitw;

static void put(bitw* w, uint32_t code, int nb) {
    w->acc = (w->acc << nb) | code; w->nacc += nb;
    while (w->nacc >= 8) {
        if (w->o >= w->cap) { w->fail = 1; w->nacc -= 8; continue; }
        w->dst[w->o++] = (uint8_t)(w->acc >> (w->nacc - 8)); w->nacc -= 8;
    }
}

/* restart > 0: an additional Clear code every `restart` input bytes (any TIFF-LZW reader accepts a Clear anywhere) */
int64_t syn_lzw_encode_restart(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, size_t restart) {
    enum { CLEAR = 256, EOI = 257, FIRST = 258, LIMIT = 4094 };
    static __thread int16_t child[4096], sib[4096];
    static __thread uint8_t ch[4096];
    bitw w = {0, 0, dst, 0, cap, 0};
    int nbits = 9, next = FIRST;
    memset(child, 0xff, sizeof child);
    put(&w, CLEAR, nbits);
    size_t pos = 0;
    while (pos < n) {
        const size_t end = restart && n - pos > restart ? pos + restart : n;
        int cur = src[pos];
        for (size_t i = pos + 1; i < end; i++) {
            uint8_t c = src[i];
            int k = child[cur];
            while (k >= 0 && ch[k] != c) k = sib[k];
            if (k >= 0) { cur = k; continue; }
            put(&w, (uint32_t)cur, nbits);
            ch[next] = c; sib[next] = child[cur]; child[cur] = (int16_t)next; child[next] = -1;
            next++;
            cur = c;
            if (next == LIMIT) {
                put(&w, CLEAR, nbits);
                memset(child, 0xff, sizeof child);
                nbits = 9; next = FIRST;
            } else if (next > (1 << nbits) - 1) {
                nbits++;
            }
        }
        put(&w, (uint32_t)cur, nbits);
        next++;
        if (next == LIMIT) { put(&w, CLEAR, nbits); nbits = 9; }
        else if (next > (1 << nbits) - 1 && nbits < 12) nbits++;
        pos = end;
        if (pos < n) {                                   /* a restart: Clear in the current width, fresh table */
            put(&w, CLEAR, nbits);
            memset(child, 0xff, sizeof child);
            nbits = 9; next = FIRST;
        }
    }
    put(&w, EOI, nbits);
    if (w.nacc) put(&w, 0, 8 - w.nacc);
    return w.fail ? -1 : (int64_t)w.o;
}

int64_t syn_lzw_encode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    return syn_lzw_encode_restart(src, n, dst, cap, 0);
}

/* horizontal differencing (predictor 2), in place, host-order words */
void syn_hdiff(uint8_t* buf, size_t rows, size_t row_samples, int spp, int bytes_per_sample) {
    for (size_t r = 0; r < rows; r++) {
        if (bytes_per_sample == 1) {
            uint8_t* p = buf + r * row_samples;
            for (size_t i = row_samples; i-- > (size_t)spp;) p[i] = (uint8_t)(p[i] - p[i - spp]);
        } else if (bytes_per_sample == 2) {
            uint16_t* p = (uint16_t*)buf + r * row_samples;
            for (size_t i = row_samples; i-- > (size_t)spp;) p[i] = (uint16_t)(p[i] - p[i - spp]);
        } else {
            uint32_t* p = (uint32_t*)buf + r * row_samples;
            for (size_t i = row_samples; i-- > (size_t)spp;) p[i] = p[i] - p[i - spp];
        }
    }
}
