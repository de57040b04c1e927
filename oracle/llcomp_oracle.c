/*
 * llcomp_oracle.c -- plain-C restatement of llcomp revision 2 (see header).
 * TEST INFRASTRUCTURE ONLY; never linked into the product library.
 *
 * Each function cites the lines of /root/reference/llcomp.hpp it follows.
 * The constant tables are rebuilt from their structure (pairs, closed forms)
 * and compared entry-by-entry with the reference arrays in tests/test_oracle.py.
 */
#include "llcomp_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- format constants (llcomp.hpp:17-32) -------------------------------- */
#define LLO_MAGIC 0x79          /* 0x77 + revision 2, llcomp.hpp:19-20 */
#define LLO_SUBSTATES 8         /* llcomp.hpp:25 */
#define LLO_NSTATES ((11 * 11 * 11 * 5 * 5 + 1) / 2 * LLO_SUBSTATES) /* :26-32 */
#define LLO_E_LIM 4             /* llcomp.hpp:22 */
#define LLO_R_LIM 6             /* llcomp.hpp:23 */
#define LLO_S_CTX 7             /* llcomp.hpp:24 */

/* ---- adaptive bit model (llcomp.hpp:250-294) ----------------------------
 * States come in (even, odd) pairs: parity is the MPS, P(bit=1)*256 of the odd
 * member is 254 minus that of the even member, and the LPS successor of pair k
 * is pair lps_pair[k] with the same parity (pair 0 swaps parity instead). */
static const uint8_t even_prob[64] = {
    123, 117, 111, 106, 101, 96, 91, 87, 83, 79, 75, 72, 68, 66, 63, 60,
    57, 54, 52, 49, 48, 45, 43, 41, 40, 38, 36, 35, 33, 32, 30, 30,
    28, 27, 26, 25, 24, 23, 22, 21, 21, 20, 19, 18, 18, 17, 17, 16,
    16, 15, 15, 14, 14, 13, 13, 13, 12, 12, 12, 11, 11, 11, 11, 7};
static const uint8_t lps_pair[64] = {
    0, 0, 1, 2, 2, 4, 4, 5, 6, 7, 8, 9, 9, 11, 11, 12,
    13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22, 23, 24,
    24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33,
    33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 38, 38, 38, 38, 39};

int llo_state_probability(int s) {
    int p = even_prob[(s >> 1) & 63];
    return (s & 1) ? 254 - p : p;
}
int llo_next_state_mps(int s) { return s < 126 ? s + 2 : s; }
int llo_next_state_lps(int s) {
    if (s < 2) return s ^ 1;
    return 2 * lps_pair[s >> 1] + (s & 1);
}

/* ---- context quantisers (llcomp.hpp:297-341), closed form --------------- */
static int clamp8(int x) { return x < -128 ? -128 : (x > 127 ? 127 : x); }
int llo_quant11(int x) {
    int v = clamp8(x), a = v < 0 ? -v : v;
    int q = (a >= 1) + (a >= 2) + (a >= 5) + (a >= 12) + (a >= 35);
    return v < 0 ? -q : q;
}
int llo_quant5(int x) {
    int v = clamp8(x), a = v < 0 ? -v : v;
    int q = (a >= 1) + (a >= 4);
    return v < 0 ? -q : q;
}

/* ---- median of three (llcomp.hpp:343-356) ------------------------------- */
int llo_median(int a, int b, int c) {
    int lo = a < b ? a : b, hi = a < b ? b : a;
    return c < lo ? lo : (c > hi ? hi : c);
}

/* ---- growable byte sink (fix D1) ---------------------------------------- */
typedef struct {
    uint8_t *p;
    size_t n, cap;
    int oom;
} sink_t;

static void sink_put(sink_t *s, int byte) {
    if (s->n == s->cap) {
        size_t nc = s->cap ? s->cap * 2 : 4096;
        uint8_t *np = (uint8_t *)realloc(s->p, nc);
        if (!np) { s->oom = 1; return; }
        s->p = np;
        s->cap = nc;
    }
    s->p[s->n++] = (uint8_t)byte;
}

/* ---- range encoder (llcomp.hpp:33-89) ----------------------------------- */
typedef struct {
    int low, range, held, pending; /* held = outstanding_byte, pending = outstanding_count */
    sink_t *out;
} renc_t;

static void renc_init(renc_t *e, sink_t *out) { /* llcomp.hpp:35 */
    e->low = 0; e->range = 0xFF00; e->held = -1; e->pending = 0; e->out = out;
}
static void renc_renorm(renc_t *e) { /* llcomp.hpp:38-58 */
    while (e->range < 0x100) {
        if (e->held < 0) {
            e->held = e->low >> 8;
        } else if (e->low <= 0xFF00) {
            sink_put(e->out, e->held);
            for (; e->pending; e->pending--) sink_put(e->out, 0xFF);
            e->held = e->low >> 8;
        } else if (e->low >= 0x10000) {
            sink_put(e->out, e->held + 1);
            for (; e->pending; e->pending--) sink_put(e->out, 0x00);
            e->held = (e->low >> 8) & 0xFF;
        } else {
            e->pending++;
        }
        e->low = (e->low & 0xFF) << 8;
        e->range <<= 8;
    }
}
static void renc_put(renc_t *e, int bit, int prob) { /* llcomp.hpp:60-73 */
    int r1 = (e->range * prob) >> 8;
    if (!bit) {
        e->range -= r1;
    } else {
        e->low += e->range - r1;
        e->range = r1;
    }
    renc_renorm(e);
}
static void renc_finish(renc_t *e) { /* llcomp.hpp:75-81 */
    e->range = 0xFF; e->low += 0xFF; renc_renorm(e);
    e->range = 0xFF; renc_renorm(e);
}

/* ---- range decoder (llcomp.hpp:91-127) ---------------------------------- */
typedef struct {
    int low, range;
    const uint8_t *p;
    size_t n, pos;
} rdec_t;

static int rdec_byte(rdec_t *d) { /* zero fill past the end, llcomp.hpp:475-479 */
    return d->pos < d->n ? d->p[d->pos++] : 0;
}
static void rdec_init(rdec_t *d, const uint8_t *p, size_t n) { /* llcomp.hpp:93-96 */
    d->p = p; d->n = n; d->pos = 0; d->range = 0xFF00;
    d->low = rdec_byte(d) << 8;
    d->low |= rdec_byte(d);
}
static int rdec_get(rdec_t *d, int prob) { /* llcomp.hpp:98-121 */
    int r1 = (d->range * prob) >> 8, bit;
    d->range -= r1;
    if (d->low < d->range) {
        bit = 0;
    } else {
        d->low -= d->range;
        d->range = r1;
        bit = 1;
    }
    if (d->range < 0x100) { /* single-step refill, llcomp.hpp:98-104 */
        d->range <<= 8;
        d->low = (d->low << 8) + rdec_byte(d);
    }
    return bit;
}

/* ---- binarisation (llcomp.hpp:166-206) ----------------------------------- */
int llo_binarize(int diff, uint8_t *cb) {
    int n = 0;
    uint32_t uv = (uint32_t)(diff < 0 ? -diff : diff);
    if (uv == 0) { cb[n++] = (0 << 1) | 1; return n; }       /* :204 */
    int e = 31 - __builtin_clz(uv);                          /* :148, :184 */
    cb[n++] = (0 << 1) | 0;                                  /* :187 */
    int ctx = 1;
    for (int i = 0; i < e; i++) {                            /* :190-192 */
        int k = ctx < LLO_E_LIM ? ctx : LLO_E_LIM; ctx++;
        cb[n++] = (uint8_t)((k << 1) | 1);
    }
    cb[n++] = (uint8_t)(((ctx < LLO_E_LIM ? ctx : LLO_E_LIM) << 1) | 0); /* :193 */
    ctx = LLO_E_LIM + 1;                                     /* :195 */
    for (int i = e - 1; i >= 0; i--) {                       /* :196-198 */
        int k = ctx < LLO_R_LIM ? ctx : LLO_R_LIM; ctx++;
        cb[n++] = (uint8_t)((k << 1) | ((uv >> i) & 1));
    }
    cb[n++] = (uint8_t)((LLO_S_CTX << 1) | (diff < 0));      /* :200-202 */
    return n;
}

/* ---- colour transform of one pixel (llcomp.hpp:396-414) ------------------ */
static void forward_planes(const uint8_t *px, int c, int16_t *dst) {
    if (c >= 3) {
        int r = px[0], g = px[1], b = px[2];
        b -= g; r -= g;
        g += (b + r) / 4;          /* C division truncates toward zero, :402 */
        dst[0] = (int16_t)r; dst[1] = (int16_t)g; dst[2] = (int16_t)b;
        for (int i = 3; i < c; i++) dst[i] = px[i];
    } else {
        for (int i = 0; i < c; i++) dst[i] = px[i];
    }
}

/* ---- neighbourhood -> (hash, predictor) (llcomp.hpp:417-430 / :494-509) -- */
typedef struct { int hash, predict; } ctx_t;

static ctx_t context_of(const int16_t *r0, const int16_t *r1, const int16_t *r2,
                        int w, int h, int width, int c, int i) {
    const int x = w * c;
    const int l  = w > 0 ? r0[x - c + i] : (h > 0 ? r1[x + i] : 128);
    const int t  = h > 0 ? r1[x + i] : l;
    const int L  = w > 1 ? r0[x - 2 * c + i] : l;
    const int tl = (h > 0 && w > 0) ? r1[x - c + i] : t;
    const int tr = (h > 0 && w < width - 1) ? r1[x + c + i] : t;
    const int T  = h > 1 ? r2[x + i] : t;
    ctx_t k;
    k.hash = llo_quant11(l - tl) + 11 * llo_quant11(tl - t) + 121 * llo_quant11(t - tr)
           + 605 * llo_quant5(L - l) + 3025 * llo_quant5(T - t);   /* :424-429 */
    k.predict = llo_median(l, l + t - tl, t);                      /* :430 */
    return k;
}

/* ---- front end: pixels -> records (llcomp.hpp:390-436) ------------------- */
typedef void (*sym_fn)(void *ud, int hash, int diff);

static int walk_tile(const uint8_t *px, size_t pitch, int w, int h, int c,
                     sym_fn fn, void *ud) {
    if (w <= 0 || h <= 0 || c <= 0) return LLO_BAD_ARG;
    const size_t stride = (size_t)w * c;
    int16_t *rows = (int16_t *)malloc(3 * stride * sizeof(int16_t));
    if (!rows) return LLO_NOMEM;
    for (int y = 0; y < h; y++) {
        int16_t *r0 = rows + (size_t)(y % 3) * stride;            /* :391-393 */
        int16_t *r1 = rows + (size_t)((y + 2) % 3) * stride;
        int16_t *r2 = rows + (size_t)((y + 1) % 3) * stride;
        const uint8_t *src = px + (size_t)y * pitch;
        for (int x = 0; x < w; x++) {
            forward_planes(src + (size_t)x * c, c, r0 + (size_t)x * c);
            for (int i = 0; i < c; i++) {
                ctx_t k = context_of(r0, r1, r2, x, y, w, c, i);
                int diff = r0[(size_t)x * c + i] - k.predict;      /* :431 */
                if (k.hash < 0) { k.hash = -k.hash; diff = -diff; } /* :433-436 */
                fn(ud, k.hash, diff);
            }
        }
    }
    free(rows);
    return LLO_OK;
}

static void store_sym(void *ud, int hash, int diff) {
    uint32_t **pp = (uint32_t **)ud;
    *(*pp)++ = LLO_SYM_PACK(hash, diff);
}
int llo_frontend_tile(const uint8_t *px, size_t pitch, int w, int h, int c, uint32_t *out) {
    uint32_t *p = out;
    return walk_tile(px, pitch, w, h, c, store_sym, &p);
}

static void count_sym(void *ud, int hash, int diff) {
    (void)hash;
    uint8_t cb[40];
    *(uint64_t *)ud += (uint64_t)llo_binarize(diff, cb);
}
uint64_t llo_count_bins(const uint8_t *px, size_t pitch, int w, int h, int c) {
    uint64_t n = 0;
    walk_tile(px, pitch, w, h, c, count_sym, &n);
    return n;
}

/* ---- coder back end: records -> bytes (llcomp.hpp:439-449) --------------- */
typedef struct {
    renc_t enc;
    uint8_t *state; /* one byte per (context, sub-state), index hash*8+ctx, :440-441 */
} coder_t;

static void code_sym(void *ud, int hash, int diff) {
    coder_t *k = (coder_t *)ud;
    uint8_t cb[40];
    int n = llo_binarize(diff, cb);
    uint8_t *row = k->state + (size_t)hash * LLO_SUBSTATES;
    for (int j = 0; j < n; j++) {
        int ctx = cb[j] >> 1, bit = cb[j] & 1;
        int s = row[ctx];
        renc_put(&k->enc, bit, llo_state_probability(s));                     /* :442 */
        row[ctx] = (uint8_t)((bit == (s & 1)) ? llo_next_state_mps(s)          /* :443, :290-292 */
                                              : llo_next_state_lps(s));
    }
}

static int sink_finish(sink_t *s, uint8_t **out, size_t *out_len) {
    if (s->oom) { free(s->p); return LLO_NOMEM; }
    if (!s->p) s->p = (uint8_t *)malloc(1);
    *out = s->p; *out_len = s->n;
    return LLO_OK;
}

int llo_encode_tile(const uint8_t *px, size_t pitch, int w, int h, int c,
                    uint8_t **out, size_t *out_len) {
    sink_t s = {0, 0, 0, 0};
    coder_t k;
    k.state = (uint8_t *)calloc(LLO_NSTATES, 1);   /* all states start at 0, :284, :385 */
    if (!k.state) return LLO_NOMEM;
    renc_init(&k.enc, &s);
    int rc = walk_tile(px, pitch, w, h, c, code_sym, &k);
    free(k.state);
    if (rc) { free(s.p); return rc; }
    renc_finish(&k.enc);                           /* :449 */
    return sink_finish(&s, out, out_len);
}

int llo_encode_symbols(const uint32_t *sym, size_t n, uint8_t **out, size_t *out_len) {
    sink_t s = {0, 0, 0, 0};
    coder_t k;
    k.state = (uint8_t *)calloc(LLO_NSTATES, 1);
    if (!k.state) return LLO_NOMEM;
    renc_init(&k.enc, &s);
    for (size_t j = 0; j < n; j++) {
        int hash = (int)(sym[j] >> 11);
        int diff = (int)(sym[j] & 0x7FF);
        if (diff & 0x400) diff -= 0x800;
        code_sym(&k, hash, diff);
    }
    free(k.state);
    renc_finish(&k.enc);
    return sink_finish(&s, out, out_len);
}

/* ---- decoder (llcomp.hpp:475-545, fix D2) -------------------------------- */
int llo_decode_tile(const uint8_t *payload, size_t len, int w, int h, int c,
                    uint8_t *px_out, size_t pitch) {
    if (w <= 0 || h <= 0 || c <= 0) return LLO_BAD_ARG;
    const size_t stride = (size_t)w * c;
    int16_t *rows = (int16_t *)calloc(3 * stride, sizeof(int16_t));
    uint8_t *state = (uint8_t *)calloc(LLO_NSTATES, 1);
    if (!rows || !state) { free(rows); free(state); return LLO_NOMEM; }
    rdec_t d;
    rdec_init(&d, payload, len);
    int rc = LLO_OK;
    for (int y = 0; y < h && !rc; y++) {
        int16_t *r0 = rows + (size_t)(y % 3) * stride;             /* :487-489 */
        int16_t *r1 = rows + (size_t)((y + 2) % 3) * stride;
        int16_t *r2 = rows + (size_t)((y + 1) % 3) * stride;
        uint8_t *dst = px_out + (size_t)y * pitch;
        for (int x = 0; x < w && !rc; x++) {
            for (int i = 0; i < c; i++) {
                ctx_t k = context_of(r0, r1, r2, x, y, w, c, i);
                int neg = 0;
                if (k.hash < 0) { k.hash = -k.hash; neg = 1; }      /* :511-515 */
                uint8_t *row = state + (size_t)k.hash * LLO_SUBSTATES;
#define GETBIN(ctx_, dst_) do { int s_ = row[ctx_];                                  \
        int b_ = rdec_get(&d, llo_state_probability(s_));                            \
        row[ctx_] = (uint8_t)((b_ == (s_ & 1)) ? llo_next_state_mps(s_)              \
                                               : llo_next_state_lps(s_));            \
        dst_ = b_; } while (0)
                int bit, diff = 0;
                GETBIN(0, bit);                                     /* :225 */
                if (!bit) {
                    int e = 0, ctx = 1;
                    int32_t value = 1;
                    for (;;) {                                       /* :230-235 */
                        int kx = ctx < LLO_E_LIM ? ctx : LLO_E_LIM; ctx++;
                        GETBIN(kx, bit);
                        if (!bit) break;
                        if (++e > 31) { rc = LLO_BAD_EXPONENT; break; }
                    }
                    if (rc) break;
                    ctx = LLO_E_LIM + 1;                             /* :237 */
                    for (int j = e - 1; j >= 0; j--) {               /* :238-240 */
                        int kx = ctx < LLO_R_LIM ? ctx : LLO_R_LIM; ctx++;
                        GETBIN(kx, bit);
                        value = (int32_t)((uint32_t)value + (uint32_t)value + (uint32_t)bit);
                    }
                    GETBIN(LLO_S_CTX, bit);                          /* :242-245 */
                    diff = bit ? -value : value;
                }
#undef GETBIN
                if (neg) diff = -diff;                               /* :526-528 */
                r0[(size_t)x * c + i] = (int16_t)(k.predict + diff); /* :529 (int16 store) */
            }
            if (rc) break;
            const int16_t *p = r0 + (size_t)x * c;
            uint8_t *q = dst + (size_t)x * c;
            if (c >= 3) {                                            /* :532-543 */
                int r = p[0], g = p[1], b = p[2];
                g -= (r + b) / 4;
                r += g; b += g;
                q[0] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
                q[1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
                q[2] = (uint8_t)(b < 0 ? 0 : (b > 255 ? 255 : b));
                for (int i = 3; i < c; i++) q[i] = (uint8_t)p[i];
            } else {                                                 /* fix D2 */
                for (int i = 0; i < c; i++) q[i] = (uint8_t)p[i];
            }
        }
    }
    free(rows);
    free(state);
    return rc;
}

/* ---- whole-image streams (llcomp.hpp:375-378, :463-470) ------------------ */
int llo_compress(const uint8_t *px, int w, int h, int c, uint8_t **out, size_t *out_len) {
    uint8_t *pay = NULL;
    size_t n = 0;
    int rc = llo_encode_tile(px, (size_t)w * c, w, h, c, &pay, &n);
    if (rc) return rc;
    uint8_t *s = (uint8_t *)malloc(n + 6);
    if (!s) { free(pay); return LLO_NOMEM; }
    s[0] = LLO_MAGIC; s[1] = (uint8_t)c;
    s[2] = (uint8_t)(w & 0xFF); s[3] = (uint8_t)((w >> 8) & 0xFF);   /* u16 truncation, D3 */
    s[4] = (uint8_t)(h & 0xFF); s[5] = (uint8_t)((h >> 8) & 0xFF);
    memcpy(s + 6, pay, n);
    free(pay);
    *out = s; *out_len = n + 6;
    return LLO_OK;
}

int llo_decompress(const uint8_t *stream, size_t len, uint8_t **px_out, int *w, int *h, int *c) {
    if (len < 6) return LLO_BAD_ARG;
    if (stream[0] != LLO_MAGIC) return LLO_BAD_MAGIC;
    *c = stream[1];
    *w = stream[2] | (stream[3] << 8);
    *h = stream[4] | (stream[5] << 8);
    size_t n = (size_t)*w * *h * *c;
    uint8_t *px = (uint8_t *)malloc(n ? n : 1);
    if (!px) return LLO_NOMEM;
    int rc = n ? llo_decode_tile(stream + 6, len - 6, *w, *h, *c, px, (size_t)*w * *c) : LLO_OK;
    if (rc) { free(px); return rc; }
    *px_out = px;
    return LLO_OK;
}

void llo_free(void *p) { free(p); }

uint64_t llo_fnv1a64(const uint8_t *p, size_t n) {
    uint64_t hsh = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) { hsh ^= p[i]; hsh *= 0x100000001b3ull; }
    return hsh;
}

/* ---- MT19937 (matches std::mt19937) + generator G ---- */
typedef struct { uint32_t mt[624]; int idx; } mt_t;
static void mt_seed(mt_t *m, uint32_t s) {
    m->mt[0] = s;
    for (int i = 1; i < 624; i++)
        m->mt[i] = 1812433253u * (m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) + (uint32_t)i;
    m->idx = 624;
}
static uint32_t mt_next(mt_t *m) {
    if (m->idx >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (m->mt[i] & 0x80000000u) | (m->mt[(i + 1) % 624] & 0x7FFFFFFFu);
            m->mt[i] = m->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1) ? 0x9908B0DFu : 0);
        }
        m->idx = 0;
    }
    uint32_t y = m->mt[m->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9D2C5680u; y ^= (y << 15) & 0xEFC60000u; y ^= y >> 18;
    return y;
}
void llo_generate(uint8_t *px, int w, int h, int c, int noise, uint32_t seed) {
    mt_t m;
    mt_seed(&m, seed);
    size_t k = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            for (int i = 0; i < c; i++) {
                if (noise < 0) { px[k++] = (uint8_t)(mt_next(&m) & 0xFF); continue; }
                int v = (x * 255 / w + y * 255 / h) / 2 + i * 10;
                if (noise > 0) v += (int)(mt_next(&m) % (uint32_t)(2 * noise + 1)) - noise;
                px[k++] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
}

/* ---- threaded batch (CPU baseline "port" leg) ---------------------------- */
typedef struct {
    const uint8_t *px; int n, w, h, c, tid, nt; uint64_t bytes; int err;
} job_t;
static void *batch_worker(void *arg) {
    job_t *j = (job_t *)arg;
    size_t img = (size_t)j->w * j->h * j->c;
    for (int k = j->tid; k < j->n; k += j->nt) {
        uint8_t *s = NULL; size_t n = 0;
        if (llo_compress(j->px + img * k, j->w, j->h, j->c, &s, &n)) { j->err = 1; return NULL; }
        j->bytes += n;
        free(s);
    }
    return NULL;
}
uint64_t llo_compress_batch_mt(const uint8_t *px, int n_images, int w, int h, int c, int nt) {
    if (nt < 1) nt = 1;
    if (nt > 256) nt = 256;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nt; t++) {
        job_t j = {px, n_images, w, h, c, t, nt, 0, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    uint64_t total = 0; int err = 0;
    for (int t = 0; t < nt; t++) { pthread_join(th[t], NULL); total += jobs[t].bytes; err |= jobs[t].err; }
    return err ? 0 : total;
}

typedef struct {
    const uint8_t *streams; const uint64_t *off; int n, tid, nt; uint64_t px; int err;
} djob_t;
static void *dbatch_worker(void *arg) {
    djob_t *j = (djob_t *)arg;
    for (int k = j->tid; k < j->n; k += j->nt) {
        uint8_t *px = NULL; int w, h, c;
        if (llo_decompress(j->streams + j->off[k], (size_t)(j->off[k + 1] - j->off[k]), &px, &w, &h, &c)) {
            j->err = 1; return NULL;
        }
        j->px += (uint64_t)w * h * c;
        free(px);
    }
    return NULL;
}
uint64_t llo_decompress_batch_mt(const uint8_t *streams, const uint64_t *offsets, int n_images, int nt) {
    if (nt < 1) nt = 1;
    if (nt > 256) nt = 256;
    pthread_t th[256];
    djob_t jobs[256];
    for (int t = 0; t < nt; t++) {
        djob_t j = {streams, offsets, n_images, t, nt, 0, 0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, dbatch_worker, &jobs[t]);
    }
    uint64_t total = 0; int err = 0;
    for (int t = 0; t < nt; t++) { pthread_join(th[t], NULL); total += jobs[t].px; err |= jobs[t].err; }
    return err ? 0 : total;
}
