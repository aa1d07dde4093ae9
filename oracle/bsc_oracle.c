/* oracle/bsc_oracle.c -- TEST INFRASTRUCTURE ONLY (see bsc_oracle.h).
 *
 * CPU restatement of the libbsc 3.3.5 hot path in plain C.  Written from the behaviour of the
 * reference (file:line cited per function), not from its text: the suffix sorter is a textbook
 * prefix-doubling sort (any correct suffix order gives the same BWT), ST-k is a
 * stable LSD byte radix sort, the QLFC coder is expressed through one generic "decision"
 * primitive over our own flat counter layout.
 */
#include "bsc_oracle.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

extern const unsigned char orc_rank_state_tab[32768 + 1];
extern const unsigned char orc_run_state_tab[8192 + 1];
extern const short orc_static_params[7][15];

static void put32(unsigned char *p, uint32_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); p[2] = (unsigned char)(v >> 16); p[3] = (unsigned char)(v >> 24); }
static uint32_t get32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static int ilog2(unsigned v) { int r = 0; while (v >>= 1) ++r; return r; }

/* ------------------------------------------------------------------------------------------ */
/* Adler-32  (adler32.cpp:83-203 computes the standard RFC-1950 checksum)                      */
/* ------------------------------------------------------------------------------------------ */
unsigned int orc_adler32(const unsigned char *p, int n)
{
    uint32_t a = 1, b = 0;
    while (n > 0) {
        int k = n < 5552 ? n : 5552; n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521u; b %= 65521u;
    }
    return (b << 16) | a;
}

/* ------------------------------------------------------------------------------------------ */
/* Forward BWT.  bwt.cpp:178-231 + libsais.c:6867-6893:                                         */
/*   SA over T[0..n) with "shorter suffix first"; p = rank of suffix 0;                         */
/*   L[0]=T[n-1]; L[j+1]=T[SA[j]-1] (j<p); L[j]=T[SA[j]-1] (j>p); returns p+1;                   */
/*   r = 2^floor(log2(n/8)); indexes[t] = rank of suffix (t+1)*r, num_indexes=(n-1)/r.          */
/* ------------------------------------------------------------------------------------------ */
static int *orc_suffix_array(const unsigned char *T, int n, int **isa_out)
{
    int *sa = malloc(sizeof(int) * (size_t)n), *rk = malloc(sizeof(int) * (size_t)n);
    int *tmp = malloc(sizeof(int) * (size_t)n), *cnt = malloc(sizeof(int) * ((size_t)n + 2));
    int *key2 = malloc(sizeof(int) * (size_t)n);
    if (!sa || !rk || !tmp || !cnt || !key2) { free(sa); free(rk); free(tmp); free(cnt); free(key2); return NULL; }

    /* depth 1: counting sort on the first byte; rank = 1 + first slot of the bucket */
    int c256[257]; memset(c256, 0, sizeof c256);
    for (int i = 0; i < n; ++i) c256[T[i] + 1]++;
    for (int c = 0; c < 256; ++c) c256[c + 1] += c256[c];
    for (int i = 0; i < n; ++i) rk[i] = 1 + c256[T[i]];
    { int pos[256]; memcpy(pos, c256, sizeof pos); for (int i = 0; i < n; ++i) sa[pos[T[i]]++] = i; }

    for (int h = 1; ; h <<= 1) {
        /* secondary key: rank of the suffix h further on; 0 when it is the empty suffix */
        /* stable LSD: first by key2, then by rk (both in [0,n]) */
        for (int i = 0; i < n; ++i) key2[i] = (i + h < n) ? rk[i + h] : 0;
        memset(cnt, 0, sizeof(int) * ((size_t)n + 2));
        for (int i = 0; i < n; ++i) cnt[key2[i] + 1]++;
        for (int v = 0; v <= n; ++v) cnt[v + 1] += cnt[v];
        for (int i = 0; i < n; ++i) tmp[cnt[key2[i]]++] = i;
        memset(cnt, 0, sizeof(int) * ((size_t)n + 2));
        for (int i = 0; i < n; ++i) cnt[rk[i] + 1]++;
        for (int v = 0; v <= n; ++v) cnt[v + 1] += cnt[v];
        for (int j = 0; j < n; ++j) { int i = tmp[j]; sa[cnt[rk[i]]++] = i; }
        /* re-rank */
        int distinct = 1; tmp[sa[0]] = 1;
        for (int j = 1; j < n; ++j) {
            int a = sa[j - 1], b = sa[j];
            int same = (rk[a] == rk[b]) && (key2[a] == key2[b]);
            if (!same) distinct++;
            tmp[b] = same ? tmp[a] : j + 1;
        }
        memcpy(rk, tmp, sizeof(int) * (size_t)n);
        if (distinct == n) break;
    }
    for (int i = 0; i < n; ++i) rk[i] -= 1;              /* 0-based ISA */
    free(tmp); free(cnt); free(key2);
    *isa_out = rk;
    return sa;
}

int orc_bwt_encode(unsigned char *T, int n, unsigned char *num_indexes, int *indexes)
{
    if (T == NULL || n < 0) return ORC_BAD_PARAMETER;
    int want_aux = (num_indexes != NULL && indexes != NULL);
    int r = 0;
    if (want_aux) {
        int mod = n / 8;
        mod |= mod >> 1; mod |= mod >> 2; mod |= mod >> 4; mod |= mod >> 8; mod |= mod >> 16; mod >>= 1;
        r = mod + 1;
        if (r < 2) return ORC_BAD_PARAMETER;            /* libsais.c:6869 rejects r < 2 */
    }
    if (n <= 1) return n;                                /* libsais.c:6845-6850 (non-aux) */

    int *isa = NULL, *sa = orc_suffix_array(T, n, &isa);
    if (!sa) return ORC_NOT_ENOUGH_MEMORY;
    unsigned char *L = malloc((size_t)n);
    if (!L) { free(sa); free(isa); return ORC_NOT_ENOUGH_MEMORY; }
    int p = isa[0];
    L[0] = T[n - 1];
    for (int j = 0; j < n; ++j) {
        if (j < p) L[j + 1] = T[sa[j] - 1];
        else if (j > p) L[j] = T[sa[j] - 1];
    }
    if (want_aux) {
        int cnt = (n - 1) / r;
        *num_indexes = (unsigned char)cnt;
        for (int t = 0; t < cnt; ++t) indexes[t] = isa[(t + 1) * r];
    }
    memcpy(T, L, (size_t)n);
    free(L); free(sa); free(isa);
    return p + 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Inverse BWT.  bwt.cpp:283-332.  Rows of the (n+1)-row matrix of T$: row 0 is "$...",        */
/* L is the last column with the '$' entry (row `index`) removed.  Walk LF from row 0.          */
/* ------------------------------------------------------------------------------------------ */
int orc_bwt_decode(unsigned char *T, int n, int index)
{
    if (T == NULL || n < 0 || index <= 0 || index > n) return ORC_BAD_PARAMETER;
    if (n <= 1) return ORC_NO_ERROR;
    uint32_t *lf = malloc(sizeof(uint32_t) * ((size_t)n + 1));
    unsigned char *out = malloc((size_t)n);
    if (!lf || !out) { free(lf); free(out); return ORC_NOT_ENOUGH_MEMORY; }
    uint32_t C[257]; memset(C, 0, sizeof C);
    for (int i = 0; i < n; ++i) C[T[i] + 1]++;
    for (int c = 0; c < 256; ++c) C[c + 1] += C[c];
    uint32_t occ[256]; memset(occ, 0, sizeof occ);
    /* full-matrix row -> index into L:  row < index: row ; row > index: row-1 ; row == index is '$' */
    for (int row = 0; row <= n; ++row) {
        if (row == index) { lf[row] = 0; continue; }
        unsigned char c = T[row < index ? row : row - 1];
        lf[row] = 1 + C[c] + occ[c]++;
    }
    uint32_t row = 0;
    for (int s = 0; s < n; ++s) {
        out[n - 1 - s] = T[row < (uint32_t)index ? row : row - 1];
        row = lf[row];
    }
    memcpy(T, out, (size_t)n);
    free(lf); free(out);
    return ORC_NO_ERROR;
}

/* ------------------------------------------------------------------------------------------ */
/* Sort Transform of order k.  st.cpp:990-1012, 200-236; st.cu:99-163 for k = 7, 8.             */
/*   P = stable sort of positions by the k bytes T[(i+d) mod n]; L[j] = T[(P[j]-1) mod n];      */
/*   returns j with P[j] == 0.                                                                  */
/* ------------------------------------------------------------------------------------------ */
int orc_st_encode(unsigned char *T, int n, int k)
{
    if (T == NULL || n < 0) return ORC_BAD_PARAMETER;
    if (k < 3 || k > 8) return ORC_BAD_PARAMETER;
    if (n <= 1) return 0;
    uint32_t *P = malloc(sizeof(uint32_t) * (size_t)n), *Q = malloc(sizeof(uint32_t) * (size_t)n);
    unsigned char *L = malloc((size_t)n);
    if (!P || !Q || !L) { free(P); free(Q); free(L); return ORC_NOT_ENOUGH_MEMORY; }
    for (int i = 0; i < n; ++i) P[i] = (uint32_t)i;
    for (int d = k - 1; d >= 0; --d) {                   /* LSD: least significant context byte first */
        uint32_t cnt[257]; memset(cnt, 0, sizeof cnt);
        for (int i = 0; i < n; ++i) cnt[T[(P[i] + (uint32_t)d) % (uint32_t)n] + 1]++;
        for (int c = 0; c < 256; ++c) cnt[c + 1] += cnt[c];
        for (int i = 0; i < n; ++i) Q[cnt[T[(P[i] + (uint32_t)d) % (uint32_t)n]]++] = P[i];
        uint32_t *t = P; P = Q; Q = t;
    }
    int index = -1;
    for (int j = 0; j < n; ++j) {
        L[j] = T[(P[j] + (uint32_t)n - 1) % (uint32_t)n];
        if (P[j] == 0) index = j;
    }
    memcpy(T, L, (size_t)n);
    free(P); free(Q); free(L);
    return index;
}

/* ------------------------------------------------------------------------------------------ */
/* Inverse sort transform.  st.cpp:1491-1527 (context boundaries 1014-1093, walk 1095-1130).     */
/*   Rows with equal k-byte context (a k-group) stand in text order.  Walking the text backwards */
/*   from a row with L = c lands in the k-group that LF (stable counting sort of L) maps it to,  */
/*   and the rows of every k-group are consumed from the last one to the first one.             */
/*   Group boundaries by induction on the context order r: row LF[i] starts an (r+1)-group iff   */
/*   row i is the first one carrying its symbol inside its r-group.                              */
/* ------------------------------------------------------------------------------------------ */
int orc_st_decode(unsigned char *T, int n, int k, int index)
{
    if (T == NULL || n < 0) return ORC_BAD_PARAMETER;
    if (index < 0 || index >= n) return ORC_BAD_PARAMETER;
    if (k < 3 || k > 8) return ORC_BAD_PARAMETER;
    if (n <= 1) return ORC_NO_ERROR;
    uint32_t *lf = malloc(sizeof(uint32_t) * (size_t)n), *grp = malloc(sizeof(uint32_t) * (size_t)n), *via = malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t *top = malloc(sizeof(uint32_t) * (size_t)n);
    unsigned char *out = malloc((size_t)n), *bucket = calloc((size_t)n, 1);
    if (!lf || !grp || !via || !top || !out || !bucket) { free(lf); free(grp); free(via); free(top); free(out); free(bucket); return ORC_NOT_ENOUGH_MEMORY; }
    uint32_t cnt[257]; memset(cnt, 0, sizeof cnt);
    for (int i = 0; i < n; ++i) cnt[T[i] + 1]++;
    for (int c = 0; c < 256; ++c) { if (cnt[c + 1]) bucket[cnt[c]] = 1; cnt[c + 1] += cnt[c]; }
    for (int i = 0; i < n; ++i) lf[i] = cnt[T[i]]++;
    for (int r = 1; r <= k; ++r) {                       /* grp[j] = first row of j's r-group */
        uint32_t g = 0;
        for (int j = 0; j < n; ++j) {
            if (bucket[j] || (r > 1 && via[j] != via[j - 1])) g = (uint32_t)j;
            grp[j] = g;
        }
        if (r < k) { for (int j = 0; j < n; ++j) top[lf[j]] = grp[j]; uint32_t *t = via; via = top; top = t; }
    }
    for (int j = 0; j < n; ++j) top[grp[j]] = (uint32_t)j; /* last row of every k-group */
    uint32_t p = (uint32_t)index;
    for (int i = n - 1; i >= 0; --i) {
        out[i] = T[p];
        uint32_t g = grp[lf[p]];
        p = top[g]; top[g] = p - 1;                      /* may wrap below the group on the very last step only */
        if (p >= (uint32_t)n) p = 0;                     /* corrupt input guard */
    }
    memcpy(T, out, (size_t)n);
    free(lf); free(grp); free(via); free(top); free(out); free(bucket);
    return ORC_NO_ERROR;
}

/* ------------------------------------------------------------------------------------------ */
/* QLFC stage 1: backward move-to-front ranks, one per run.  qlfc.cpp:398-455.                  */
/* ------------------------------------------------------------------------------------------ */
int orc_qlfc_transform(const unsigned char *in, int n, unsigned char *ranks, unsigned char mtf[256])
{
    unsigned char seen[256]; memset(seen, 0, sizeof seen);
    for (int c = 0; c < 256; ++c) mtf[c] = (unsigned char)c;
    if (in[n - 1] == 0) { mtf[0] = 1; mtf[1] = 0; }      /* keep "front != first run's symbol" */

    int w = n, nsym = 0;
    for (int i = n - 1; i >= 0; ) {
        unsigned char c = in[i];
        while (i >= 0 && in[i] == c) --i;
        /* pull c to the front, remembering how deep it was */
        int depth = 1; unsigned char moving = mtf[0]; mtf[0] = c;
        for (;; ++depth) { unsigned char t = mtf[depth]; mtf[depth] = moving; if (t == c) break; moving = t; }
        if (!seen[c]) { seen[c] = 1; depth = nsym++; }   /* last occurrence: ordinal from the end */
        ranks[--w] = (unsigned char)depth;
    }
    ranks[n - 1] = 1;                                    /* qlfc.cpp:444 */
    for (int d = 1; d < 256; ++d)
        if (!seen[mtf[d]]) { mtf[d] = mtf[d - 1]; break; }
    int R = n - w;
    memmove(ranks, ranks + w, (size_t)R);
    return R;
}

/* ------------------------------------------------------------------------------------------ */
/* Binary range coder.  coder/common/rangecoder.h:38-271 (32-bit range, 16-bit output units).   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t low32, carry, range, cache, pending;
    unsigned char *start; long pos, eob;
} rc_enc;

static void rc_put16(rc_enc *e, uint32_t v) { e->start[e->pos] = (unsigned char)v; e->start[e->pos + 1] = (unsigned char)(v >> 8); e->pos += 2; }

static void rc_shift(rc_enc *e)                          /* rangecoder.h:83-114 */
{
    if (e->low32 < 0xffff0000u || e->carry) {
        rc_put16(e, e->cache + e->carry);
        for (; e->pending; --e->pending) rc_put16(e, e->carry - 1);
        e->cache = e->low32 >> 16; e->carry = 0;
    } else e->pending++;
    e->low32 <<= 16;
}

static void rc_enc_init(rc_enc *e, unsigned char *out, int outSize)
{
    e->low32 = 0; e->carry = 0; e->range = 0xffffffffu; e->cache = 0; e->pending = 0;
    e->start = out; e->pos = 0; e->eob = (long)outSize - 16;
}

static void rc_encode(rc_enc *e, unsigned bit, int p)   /* rangecoder.h:145-177, P = 12 */
{
    if (e->range < 0x10000u) { rc_shift(e); e->range <<= 16; }
    uint32_t r = (e->range >> 12) * (uint32_t)p;
    if (bit) {
        uint32_t s = e->low32 + r; if (s < e->low32) e->carry++; e->low32 = s;
        e->range -= r;
    } else e->range = r;
}

static int rc_enc_finish(rc_enc *e)                      /* rangecoder.h:129-143 */
{
    if (e->range < 0x10000u) rc_shift(e);
    rc_shift(e); rc_shift(e); rc_shift(e);
    return (int)e->pos;
}

typedef struct { const unsigned char *in; uint32_t code, range; } rc_dec;

static uint32_t rc_get16(rc_dec *d) { uint32_t v = (uint32_t)d->in[0] | ((uint32_t)d->in[1] << 8); d->in += 2; return v; }

static void rc_dec_init(rc_dec *d, const unsigned char *in)  /* rangecoder.h:203-211 */
{
    d->in = in; d->code = 0; d->range = 0xffffffffu;
    for (int i = 0; i < 3; ++i) d->code = (d->code << 16) | rc_get16(d);
}

static unsigned rc_decode(rc_dec *d, int p)              /* rangecoder.h:224-240 */
{
    if (d->range < 0x10000u) { d->range <<= 16; d->code = (d->code << 16) | rc_get16(d); }
    uint32_t r = (d->range >> 12) * (uint32_t)p;
    if (d->code >= r) { d->code -= r; d->range -= r; return 1; }
    d->range = r; return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Static QLFC model: every binary decision mixes three 12-bit counters (per-symbol,            */
/* per-state, shared) with fixed weights /32, then moves each counter towards the bit.          */
/* qlfc_model.h:178-241 (which counters exist), predictor.h:45-61 (update rule),                */
/* qlfc.cpp:949-1125 (which decision uses which counters).                                      */
/* ------------------------------------------------------------------------------------------ */
enum { K_RANK_T, K_RANK_E, K_RANK_M, K_RANK_P, K_RUN_T, K_RUN_E, K_RUN_M };

typedef struct { short shared[256], by_state[256][256], by_char[256][256]; } wide_bank;
typedef struct { short shared[32], by_state[256][32], by_char[256][32]; } narrow_bank;

typedef struct {
    short rt_shared, rt_state[256], rt_char[256];
    short re_shared[8], re_state[256][8], re_char[256][8];
    wide_bank rm[8], rp;
    short ut_shared, ut_state[256], ut_char[256];
    narrow_bank ue, um[32];
} orc_model;

static orc_model *model_new(void)
{
    orc_model *m = malloc(sizeof *m);
    if (m) { short *p = (short *)m; for (size_t i = 0; i < sizeof *m / sizeof(short); ++i) p[i] = 2048; }  /* qlfc_model.cpp:70-71 */
    return m;
}

/* predictor.h:55-61 / 45-51 (both spellings give the same integers) */
static void ctr_move(short *x, const short *q, unsigned bit)
{
    int p = *x;
    if (bit) p -= ((p - q[2]) * q[3]) >> 12;
    else     p += ((4096 - q[0] - p) * q[1]) >> 12;
    *x = (short)p;
}

static int decide_p(int k, const short *s, const short *c, const short *g)
{
    const short *w = orc_static_params[k];
    return ((*c) * w[0] + (*s) * w[1] + (*g) * w[2]) >> 5;
}

static void decide_learn(int k, short *s, short *c, short *g, unsigned bit)
{
    const short *w = orc_static_params[k];
    ctr_move(s, w + 3, bit); ctr_move(c, w + 7, bit); ctr_move(g, w + 11, bit);
}

/* ---- adaptive variant (coder id 2, `-e2`): same decisions and the same three counters, but their probabilities go   */
/* through a per-context logistic mixer with on-line weights plus a 17-point probability map (predictor.h:74-213,      */
/* qlfc_model.h:38-113 M_* constants, 225-231 mixer arrays, qlfc_model.cpp:49-68 start values).                         */
extern const unsigned char orc_stretch_le16[], orc_squash_le16[];
extern const short orc_adaptive_params[7][19];
static int tab16(const unsigned char *t, int i) { return (int16_t)(uint16_t)(t[2 * i] | (t[2 * i + 1] << 8)); }
static int stretch(int p) { return tab16(orc_stretch_le16, p); }              /* tables.h:1848 */
static int squash(int s)  { return tab16(orc_squash_le16, 2048 + s); }        /* tables.h:1853 */

typedef struct { int16_t s0, s1, s2; int mixed, index; short map[17]; int w0, w1, w2; } orc_mixer;
typedef struct { orc_mixer rank[256], rankExp[8][8], rankMan[8], rankEsc[256], run[256], runExp[32][32], runMan[32]; } orc_mixers;

static void mixer_init(orc_mixer *m)                                           /* predictor.h:96-103 */
{
    m->w0 = m->w1 = 2048 << 5; m->w2 = 0;
    for (int p = 0; p < 17; ++p) m->map[p] = (short)squash((p - 8) * 256);
}
static orc_mixers *mixers_new(void)
{
    orc_mixers *x = calloc(1, sizeof *x);
    if (x) { orc_mixer *m = (orc_mixer *)x; for (size_t i = 0; i < sizeof *x / sizeof *m; ++i) mixer_init(m + i); }
    return x;
}
static int mixer_mixup(orc_mixer *m, int pChar, int pState, int pShared)       /* predictor.h:105-123: arguments in this order */
{
    m->s0 = (int16_t)stretch(pChar); m->s1 = (int16_t)stretch(pState); m->s2 = (int16_t)stretch(pShared);
    int16_t sp = (int16_t)((m->s0 * m->w0 + m->s1 * m->w1 + m->s2 * m->w2) >> 17);   /* the reference keeps this in a short */
    if (sp < -2047) sp = -2047;
    if (sp > 2047) sp = 2047;
    m->index = (sp + 2048) >> 8;
    int weight = sp & 255, probability = squash(sp);
    int mapped = m->map[m->index] + (((m->map[m->index + 1] - m->map[m->index]) * weight) >> 8);
    return m->mixed = (3 * probability + mapped) >> 2;
}
static void mixer_learn(orc_mixer *m, const short *q, unsigned bit)            /* predictor.h:185-211; q = {TH0,AR0,TH1,AR1,LR0,LR1,LR2} */
{
    ctr_move(&m->map[m->index], q, bit); ctr_move(&m->map[m->index + 1], q, bit);
    int eps = m->mixed - (bit ? 1 : 4095);
    /* int products as the reference computes them (two's-complement wrap on overflow) */
    m->w0 -= (int)((uint32_t)(q[4] * eps) * (uint32_t)(int)m->s0) >> 16;
    m->w1 -= (int)((uint32_t)(q[5] * eps) * (uint32_t)(int)m->s1) >> 16;
    m->w2 -= (int)((uint32_t)(q[6] * eps) * (uint32_t)(int)m->s2) >> 16;
}

typedef struct {
    int ctxRank0, ctxRank4, ctxRun, maxRank, avgRank;
    unsigned char rankHist[256], runHist[256];
} run_ctx;

static void run_ctx_init(run_ctx *x) { memset(x, 0, sizeof *x); x->maxRank = 7; }

static int rank_state(const run_ctx *x, int c) { return orc_rank_state_tab[(x->ctxRun << 11) | (x->ctxRank4 << 3) | x->rankHist[c]]; }
static int run_state(const run_ctx *x, int c, int rank0)
{
    int h = x->runHist[c];
    return orc_run_state_tab[(x->ctxRank0 << 10) | (x->ctxRun << 6) | ((rank0 < 7 ? rank0 : 7) << 3) | (h < 7 ? h : 7)];
}
static void run_ctx_slide(run_ctx *x, int rank0, int run)   /* qlfc.cpp:1123-1125 */
{
    x->ctxRank0 = ((x->ctxRank0 << 1) | (rank0 == 0)) & 0x7;
    x->ctxRank4 = ((x->ctxRank4 << 2) | (rank0 < 3 ? rank0 : 3)) & 0xff;
    x->ctxRun   = ((x->ctxRun << 1) | (run < 3)) & 0xf;
}

/* which of the 256 symbols can still appear at this point of the MTF-order header (qlfc.cpp:857-891) */
static void header_options(const unsigned char *used, int prev, int prefix, int bit, int *can0, int *can1)
{
    *can0 = *can1 = 0;
    for (int c = 0; c < 256; ++c)
        if ((c == prev || !used[c]) && (c >> (bit + 1)) == prefix) { if (c & (1 << bit)) *can1 = 1; else *can0 = 1; }
}

/* one decision of either model-1 coder: static (mx == NULL) mixes with fixed weights, adaptive goes through mixer MIX */
static int decide2_p(const orc_mixers *mx, int k, orc_mixer *mix, const short *s, const short *c, const short *g)
{
    return mx ? mixer_mixup(mix, *c, *s, *g) : decide_p(k, s, c, g);
}
static void decide2_learn(const orc_mixers *mx, int k, orc_mixer *mix, short *s, short *c, short *g, unsigned bit)
{
    if (!mx) { decide_learn(k, s, c, g, bit); return; }
    const short *w = orc_adaptive_params[k];
    ctr_move(s, w, bit); ctr_move(c, w + 4, bit); ctr_move(g, w + 8, bit);
    mixer_learn(mix, w + 12, bit);
}
#define ENC(K, S, C, G, MIX, BIT) do { short *s_ = (S), *c_ = (C), *g_ = (G); unsigned b_ = (BIT); orc_mixer *m_ = mx ? (MIX) : NULL; \
        int p_ = decide2_p(mx, (K), m_, s_, c_, g_); decide2_learn(mx, (K), m_, s_, c_, g_, b_); rc_encode(&rc, b_, p_); } while (0)
#define MAXI(a, b) ((a) > (b) ? (a) : (b))

static int qlfc1_encode_block(const unsigned char *in, unsigned char *out, int inSize, int outSize, int adaptive)
{
    orc_model *m = model_new();
    orc_mixers *mx = adaptive ? mixers_new() : NULL;
    unsigned char *ranks = malloc((size_t)inSize > 0 ? (size_t)inSize : 1);
    if (!m || !ranks || (adaptive && !mx)) { free(m); free(ranks); free(mx); return ORC_NOT_ENOUGH_MEMORY; }
    unsigned char mtf[256];
    int R = orc_qlfc_transform(in, inSize, ranks, mtf);

    run_ctx x; run_ctx_init(&x);
    rc_enc rc; rc_enc_init(&rc, out, outSize);
    for (int b = 31; b >= 0; --b) rc_encode(&rc, ((unsigned)inSize >> b) & 1, 2048);

    unsigned char used[256]; memset(used, 0, sizeof used);
    int prev = -1;
    for (int d = 0; d < 256; ++d) {
        int c = mtf[d];
        for (int bit = 7; bit >= 0; --bit) {
            int can0, can1; header_options(used, prev, c >> (bit + 1), bit, &can0, &can1);
            if (can0 && can1) rc_encode(&rc, (c >> bit) & 1, 2048);
        }
        if (c == prev) { x.maxRank = ilog2((unsigned)(d - 1)); break; }
        prev = c; used[c] = 1;
    }

    int pos = 0, result = 0;
    for (int t = 0; t < R; ++t) {
        if (rc.pos >= rc.eob) { result = ORC_NOT_COMPRESSIBLE; break; }   /* qlfc.cpp:898-901 */
        int c = in[pos], run = 1;
        while (pos + run < inSize && in[pos + run] == c) ++run;
        pos += run;
        int rank = ranks[t];
        int st = rank_state(&x, c), h = x.rankHist[c];

        if (x.avgRank < 32) {
            ENC(K_RANK_T, &m->rt_state[st], &m->rt_char[c], &m->rt_shared, &mx->rank[c], rank != 1);
            if (rank == 1) x.rankHist[c] = 0;
            else {
                int e = ilog2((unsigned)rank); x.rankHist[c] = (unsigned char)e;
                for (int b = 1; b < e; ++b) ENC(K_RANK_E, &m->re_state[st][b - 1], &m->re_char[c][b - 1], &m->re_shared[b - 1], &mx->rankExp[MAXI(h, b)][b], 1);
                if (e < x.maxRank)          ENC(K_RANK_E, &m->re_state[st][e - 1], &m->re_char[c][e - 1], &m->re_shared[e - 1], &mx->rankExp[MAXI(h, e)][e], 0);
                wide_bank *bk = &m->rm[e];
                for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                    unsigned b = ((unsigned)rank >> bit) & 1;
                    ENC(K_RANK_M, &bk->by_state[st][node], &bk->by_char[c][node], &bk->shared[node], &mx->rankMan[e], b);
                    node = 2 * node + (int)b;
                }
            }
        } else {
            x.rankHist[c] = (unsigned char)ilog2((unsigned)rank);
            for (int node = 1, bit = x.maxRank; bit >= 0; --bit) {
                unsigned b = ((unsigned)rank >> bit) & 1;
                ENC(K_RANK_P, &m->rp.by_state[st][node], &m->rp.by_char[c][node], &m->rp.shared[node], &mx->rankEsc[node], b);
                node = 2 * node + (int)b;
            }
        }
        x.avgRank = (x.avgRank * 124 + rank * 4) >> 7;
        int rank0 = rank - 1;
        st = run_state(&x, c, rank0); h = x.runHist[c];

        ENC(K_RUN_T, &m->ut_state[st], &m->ut_char[c], &m->ut_shared, &mx->run[c], run != 1);
        if (run == 1) x.runHist[c] = (unsigned char)((x.runHist[c] + 2) >> 2);
        else {
            int e = ilog2((unsigned)run); x.runHist[c] = (unsigned char)((x.runHist[c] + 3 * e + 3) >> 2);
            for (int b = 1; b < e; ++b) ENC(K_RUN_E, &m->ue.by_state[st][b - 1], &m->ue.by_char[c][b - 1], &m->ue.shared[b - 1], &mx->runExp[MAXI(h, b)][b], 1);
            ENC(K_RUN_E, &m->ue.by_state[st][e - 1], &m->ue.by_char[c][e - 1], &m->ue.shared[e - 1], &mx->runExp[MAXI(h, e)][e], 0);
            narrow_bank *bk = &m->um[e];
            for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                unsigned b = ((unsigned)run >> bit) & 1;
                ENC(K_RUN_M, &bk->by_state[st][node], &bk->by_char[c][node], &bk->shared[node], &mx->runMan[e], b);
                node = (e <= 5) ? 2 * node + (int)b : node + 1;      /* qlfc.cpp:1119 */
            }
        }
        run_ctx_slide(&x, rank0, run);
    }
    if (result == 0) result = rc_enc_finish(&rc);
    free(m); free(ranks); free(mx);
    return result;
}
int orc_qlfc_static_encode_block(const unsigned char *in, unsigned char *out, int inSize, int outSize) { return qlfc1_encode_block(in, out, inSize, outSize, 0); }
int orc_qlfc_adaptive_encode_block(const unsigned char *in, unsigned char *out, int inSize, int outSize) { return qlfc1_encode_block(in, out, inSize, outSize, 1); }

#define DEC(K, S, C, G, MIX, BITVAR) do { short *s_ = (S), *c_ = (C), *g_ = (G); orc_mixer *m_ = mx ? (MIX) : NULL; \
        (BITVAR) = rc_decode(&rc, decide2_p(mx, (K), m_, s_, c_, g_)); decide2_learn(mx, (K), m_, s_, c_, g_, (BITVAR)); } while (0)

static int qlfc1_decode_block(const unsigned char *in, unsigned char *out, int adaptive)
{
    orc_model *m = model_new();
    orc_mixers *mx = adaptive ? mixers_new() : NULL;
    if (!m || (adaptive && !mx)) { free(m); free(mx); return ORC_NOT_ENOUGH_MEMORY; }
    run_ctx x; run_ctx_init(&x);
    rc_dec rc; rc_dec_init(&rc, in);
    uint32_t n32 = 0; for (int b = 0; b < 32; ++b) n32 = (n32 << 1) | rc_decode(&rc, 2048);
    int n = (int)n32;

    unsigned char mtf[256], used[256]; memset(used, 0, sizeof used); memset(mtf, 0, sizeof mtf);
    int prev = -1;
    for (int d = 0; d < 256; ++d) {
        int c = 0;
        for (int bit = 7; bit >= 0; --bit) {
            int can0, can1; header_options(used, prev, c, bit, &can0, &can1);
            if (can0 && can1) c = 2 * c + (int)rc_decode(&rc, 2048);
            else if (can1) c = 2 * c + 1;
            else if (can0) c = 2 * c;
            /* neither: reference leaves c unshifted (qlfc.cpp:1719-1723) */
        }
        mtf[d] = (unsigned char)c;
        if (c == prev) { x.maxRank = ilog2((unsigned)(d - 1)); break; }
        prev = c; used[c] = 1;
    }

    for (int i = 0; i < n; ) {
        int c = mtf[0], rank = 1; unsigned b;
        int st = rank_state(&x, c), h = x.rankHist[c];
        if (x.avgRank < 32) {
            DEC(K_RANK_T, &m->rt_state[st], &m->rt_char[c], &m->rt_shared, &mx->rank[c], b);
            if (!b) x.rankHist[c] = 0;
            else {
                int e = 1;
                while (e != x.maxRank) {
                    DEC(K_RANK_E, &m->re_state[st][e - 1], &m->re_char[c][e - 1], &m->re_shared[e - 1], &mx->rankExp[MAXI(h, e) & 7][e & 7], b);
                    if (!b) break;
                    ++e;
                }
                x.rankHist[c] = (unsigned char)e;
                wide_bank *bk = &m->rm[e];
                for (int bit = e - 1; bit >= 0; --bit) {
                    DEC(K_RANK_M, &bk->by_state[st][rank], &bk->by_char[c][rank], &bk->shared[rank], &mx->rankMan[e & 7], b);
                    rank = 2 * rank + (int)b;
                }
            }
        } else {
            rank = 0;
            for (int node = 1, bit = x.maxRank; bit >= 0; --bit) {
                DEC(K_RANK_P, &m->rp.by_state[st][node], &m->rp.by_char[c][node], &m->rp.shared[node], &mx->rankEsc[node & 255], b);
                node = 2 * node + (int)b; rank = 2 * rank + (int)b;
            }
            x.rankHist[c] = (unsigned char)ilog2((unsigned)rank);
        }
        /* push the current symbol `rank` places back (qlfc.cpp:1830-1860) */
        for (int r = 0; r < rank; ++r) mtf[r] = mtf[r + 1];
        mtf[rank] = (unsigned char)c;

        x.avgRank = (x.avgRank * 124 + rank * 4) >> 7;
        int rank0 = rank - 1;
        st = run_state(&x, c, rank0); h = x.runHist[c];
        int run = 1;
        DEC(K_RUN_T, &m->ut_state[st], &m->ut_char[c], &m->ut_shared, &mx->run[c], b);
        if (!b) x.runHist[c] = (unsigned char)((x.runHist[c] + 2) >> 2);
        else {
            int e = 1;
            for (;;) {
                DEC(K_RUN_E, &m->ue.by_state[st][e - 1], &m->ue.by_char[c][e - 1], &m->ue.shared[e - 1], &mx->runExp[MAXI(h, e) & 31][e & 31], b);
                if (!b) break;
                ++e;
            }
            x.runHist[c] = (unsigned char)((x.runHist[c] + 3 * e + 3) >> 2);
            narrow_bank *bk = &m->um[e];
            for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                DEC(K_RUN_M, &bk->by_state[st][node], &bk->by_char[c][node], &bk->shared[node], &mx->runMan[e & 31], b);
                run = 2 * run + (int)b;
                node = (e <= 5) ? 2 * node + (int)b : node + 1;
            }
        }
        run_ctx_slide(&x, rank0, run);
        for (; run > 0; --run) out[i++] = (unsigned char)c;
    }
    free(m); free(mx);
    return n;
}
int orc_qlfc_static_decode_block(const unsigned char *in, unsigned char *out) { return qlfc1_decode_block(in, out, 0); }
int orc_qlfc_adaptive_decode_block(const unsigned char *in, unsigned char *out) { return qlfc1_decode_block(in, out, 1); }

/* ------------------------------------------------------------------------------------------ */
/* Fast QLFC coder (coder id 3, `-e0`).  qlfc.cpp:1135-1336 (encoder), 1933-2127 (decoder),      */
/* qlfc_model.h:243-259 (QlfcStatisticalModel2), qlfc_model.cpp:73-74 (start values),            */
/* predictor.h:63-71 (shift updates), rangecoder.h:145-177, 213-240 (precision template P).      */
/* One counter per decision, selected by the run's symbol and the position in the code only.    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    short re[256][8], rm[256][8][256];          /* rank: exponent / mantissa-tree counters, 13-bit probabilities */
    short ue[256][32], um[256][32][32];         /* run length: exponent / mantissa counters, 11-bit probabilities */
} orc_fast_model;

/* what one decision kind does: precision of the coder step, speed of the counter, its two targets */
typedef struct { int P, R, to0, to1; } fast_kind;
static const fast_kind FK_RANK_T = {13, 4, 8016, 83}, FK_RANK_E = {13, 4, 8114, 122}, FK_RANK_M = {13, 7, 7999, 235};
static const fast_kind FK_RUN_T = {11, 5, 2025, 42}, FK_RUN_E = {11, 4, 1962, 142}, FK_RUN_M = {11, 6, 1951, 147}, FK_RUN_L = {11, 5, 1987, 46};

static orc_fast_model *fast_model_new(void)
{
    orc_fast_model *m = malloc(sizeof *m);
    if (!m) return NULL;
    short *p = (short *)m->re; for (size_t i = 0; i < (sizeof m->re + sizeof m->rm) / sizeof(short); ++i) p[i] = 4096;   /* re and rm are contiguous */
    p = (short *)m->ue; for (size_t i = 0; i < (sizeof m->ue + sizeof m->um) / sizeof(short); ++i) p[i] = 1024;
    return m;
}

static void fast_learn(short *x, const fast_kind *k, unsigned bit) { int p = *x; *x = (short)(p - ((p - (bit ? k->to1 : k->to0)) >> k->R)); }

static void rc_add_low(rc_enc *e, uint32_t add) { uint32_t s = e->low32 + add; if (s < e->low32) e->carry++; e->low32 = s; }

static void rc_encode_p(rc_enc *e, unsigned bit, int p, int P)
{
    if (e->range < 0x10000u) { rc_shift(e); e->range <<= 16; }
    uint32_t r = (e->range >> P) * (uint32_t)p;
    if (bit) { rc_add_low(e, r); e->range -= r; } else e->range = r;
}

/* rangecoder.h:165-177 called with an UNNORMALISED bit value (qlfc.cpp:1174 passes `c & (1 << bit)`): the    */
/* reference uses (0 - bitval) as a mask, so for bitval = 2^k the low k bits of the addend are cleared.      */
static void rc_encode_half_masked(rc_enc *e, uint32_t bitval)
{
    if (e->range < 0x10000u) { rc_shift(e); e->range <<= 16; }
    uint32_t r = e->range >> 1, m = 0u - bitval;
    rc_add_low(e, m & r);
    e->range = r + (m & (e->range - r - r));
}

static unsigned rc_decode_p(rc_dec *d, int p, int P)
{
    if (d->range < 0x10000u) { d->range <<= 16; d->code = (d->code << 16) | rc_get16(d); }
    uint32_t r = (d->range >> P) * (uint32_t)p;
    if (d->code >= r) { d->code -= r; d->range -= r; return 1; }
    d->range = r; return 0;
}

#define FENC(KIND, CTR, BIT) do { short *x_ = (CTR); unsigned b_ = (BIT); int p_ = *x_; fast_learn(x_, &(KIND), b_); rc_encode_p(&rc, b_, p_, (KIND).P); } while (0)
#define FDEC(KIND, CTR, BITVAR) do { short *x_ = (CTR); (BITVAR) = rc_decode_p(&rc, *x_, (KIND).P); fast_learn(x_, &(KIND), (BITVAR)); } while (0)

int orc_qlfc_fast_encode_block(const unsigned char *in, unsigned char *out, int inSize, int outSize)
{
    orc_fast_model *m = fast_model_new();
    unsigned char *ranks = malloc((size_t)inSize > 0 ? (size_t)inSize : 1);
    if (!m || !ranks) { free(m); free(ranks); return ORC_NOT_ENOUGH_MEMORY; }
    unsigned char mtf[256];
    int R = orc_qlfc_transform(in, inSize, ranks, mtf);

    rc_enc rc; rc_enc_init(&rc, out, outSize);
    for (int b = 31; b >= 0; --b) rc_encode(&rc, ((unsigned)inSize >> b) & 1, 2048);

    unsigned char used[256]; memset(used, 0, sizeof used);
    int prev = -1;
    for (int d = 0; d < 256; ++d) {
        int c = mtf[d];
        for (int bit = 7; bit >= 0; --bit) {
            int can0, can1; header_options(used, prev, c >> (bit + 1), bit, &can0, &can1);
            if (can0 && can1) rc_encode_half_masked(&rc, (uint32_t)(c & (1 << bit)));
        }
        if (c == prev) break;
        prev = c; used[c] = 1;
    }

    int pos = 0, result = 0;
    for (int t = 0; t < R; ++t) {
        if (rc.pos >= rc.eob) { result = ORC_NOT_COMPRESSIBLE; break; }   /* qlfc.cpp:1192-1195 */
        int c = in[pos], run = 1;
        while (pos + run < inSize && in[pos + run] == c) ++run;
        pos += run;
        int rank = ranks[t];

        FENC(FK_RANK_T, &m->re[c][0], rank != 1);
        if (rank != 1) {
            int e = ilog2((unsigned)rank);
            for (int b = 1; b < e; ++b) FENC(FK_RANK_E, &m->re[c][b], 1);
            if (e < 7)                  FENC(FK_RANK_E, &m->re[c][e], 0);
            for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                unsigned b = ((unsigned)rank >> bit) & 1;
                FENC(FK_RANK_M, &m->rm[c][e][node], b);
                node = 2 * node + (int)b;
            }
        }
        FENC(FK_RUN_T, &m->ue[c][0], run != 1);
        if (run != 1) {
            int e = ilog2((unsigned)run);
            for (int b = 1; b < e; ++b) FENC(FK_RUN_E, &m->ue[c][b], 1);
            FENC(FK_RUN_E, &m->ue[c][e], 0);
            for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                unsigned b = ((unsigned)run >> bit) & 1;
                if (e <= 5) { FENC(FK_RUN_M, &m->um[c][e][node], b); node = 2 * node + (int)b; }
                else        { FENC(FK_RUN_L, &m->um[c][e][node], b); node = node + 1; }
            }
        }
    }
    if (result == 0) result = rc_enc_finish(&rc);
    free(m); free(ranks);
    return result;
}

int orc_qlfc_fast_decode_block(const unsigned char *in, unsigned char *out)
{
    orc_fast_model *m = fast_model_new();
    if (!m) return ORC_NOT_ENOUGH_MEMORY;
    rc_dec rc; rc_dec_init(&rc, in);
    uint32_t n32 = 0; for (int b = 0; b < 32; ++b) n32 = (n32 << 1) | rc_decode(&rc, 2048);
    int n = (int)n32;

    unsigned char mtf[256], used[256]; memset(used, 0, sizeof used); memset(mtf, 0, sizeof mtf);
    int prev = -1;
    for (int d = 0; d < 256; ++d) {
        int c = 0;
        for (int bit = 7; bit >= 0; --bit) {
            int can0, can1; header_options(used, prev, c, bit, &can0, &can1);
            if (can0 && can1) c = 2 * c + (int)rc_decode_p(&rc, 1, 1);
            else if (can1) c = 2 * c + 1;
            else if (can0) c = 2 * c;
        }
        mtf[d] = (unsigned char)c;
        if (c == prev) break;
        prev = c; used[c] = 1;
    }

    for (int i = 0; i < n; ) {
        int c = mtf[0], rank = 1; unsigned b;
        FDEC(FK_RANK_T, &m->re[c][0], b);
        if (b) {
            int e = 1;
            while (e < 7) { FDEC(FK_RANK_E, &m->re[c][e], b); if (!b) break; ++e; }
            for (int bit = e - 1; bit >= 0; --bit) { FDEC(FK_RANK_M, &m->rm[c][e][rank], b); rank = 2 * rank + (int)b; }
        }
        for (int r = 0; r < rank; ++r) mtf[r] = mtf[r + 1];
        mtf[rank] = (unsigned char)c;

        int run = 1;
        FDEC(FK_RUN_T, &m->ue[c][0], b);
        if (b) {
            int e = 1;
            for (;;) { FDEC(FK_RUN_E, &m->ue[c][e], b); if (!b) break; if (++e >= 31) break; }   /* 31: guard against corrupt input */
            for (int node = 1, bit = e - 1; bit >= 0; --bit) {
                if (e <= 5) { FDEC(FK_RUN_M, &m->um[c][e][node], b); node = 2 * node + (int)b; }
                else        { FDEC(FK_RUN_L, &m->um[c][e][node], b); node = node + 1; }
                run = 2 * run + (int)b;
            }
        }
        for (; run > 0 && i < n; --run) out[i++] = (unsigned char)c;
    }
    free(m);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* Coder container.  coder.cpp:52-59, 70-109, 111-155 (serial), 159-240 (parallel), 273-347.    */
/* ------------------------------------------------------------------------------------------ */
int orc_coder_num_blocks(int n)
{
    if (n < 256 * 1024) return 1;
    if (n < 4 * 1024 * 1024) return 2;
    if (n < 16 * 1024 * 1024) return 4;
    return 8;
}

void orc_coder_split_blocks(const unsigned char *in, int n, int nBlocks, int *start, int *size)
{
    int changes = 0;
    for (int i = 1; i < n; i += 32) changes += (in[i] != in[i - 1]);
    if (changes > nBlocks) {
        int per = changes / nBlocks, seen = 0, id = 0;
        start[0] = 0;
        for (int i = 1; i < n && id < nBlocks - 1; i += 32) {
            if (in[i] != in[i - 1] && ++seen == per) { seen = 0; size[id] = i - start[id]; start[++id] = i; }
        }
        size[nBlocks - 1] = n - start[nBlocks - 1];
    } else {
        for (int p = 0; p < nBlocks; ++p) {
            start[p] = (n / nBlocks) * p;
            size[p] = (p != nBlocks - 1) ? n / nBlocks : n - (n / nBlocks) * (nBlocks - 1);
        }
    }
}

static int encode_block(int coder, const unsigned char *in, unsigned char *out, int inSize, int outSize)   /* coder.cpp:61-68 */
{
    return coder == 3 ? orc_qlfc_fast_encode_block(in, out, inSize, outSize) : qlfc1_encode_block(in, out, inSize, outSize, coder == 2);
}
static int decode_block(int coder, const unsigned char *in, unsigned char *out)                             /* coder.cpp:264-271 */
{
    return coder == 3 ? orc_qlfc_fast_decode_block(in, out) : qlfc1_decode_block(in, out, coder == 2);
}

int orc_coder_compress(const unsigned char *in, unsigned char *out, int n, int coder, int features)
{
    if (coder < 1 || coder > 3) return ORC_BAD_PARAMETER;
    int nBlocks = orc_coder_num_blocks(n);
    if (nBlocks == 1) {
        int r = encode_block(coder, in, out + 1, n, n - 1);
        if (r >= 0) { out[0] = 1; r += 1; }
        return r;
    }
    int start[8], size[8];
    orc_coder_split_blocks(in, n, nBlocks, start, size);
    out[0] = (unsigned char)nBlocks;
    int ptr = 1 + 8 * nBlocks;
    if (features & ORC_FEATURE_MULTITHREADING) {         /* coder.cpp:159-240 */
        unsigned char *tmp = malloc((size_t)n + 256);
        int res[8], total = ptr;
        if (!tmp) return ORC_NOT_ENOUGH_MEMORY;
        for (int b = 0; b < nBlocks; ++b) {
            res[b] = encode_block(coder, in + start[b], tmp + start[b], size[b], size[b]);
            if (res[b] < 0) res[b] = size[b];
            total += res[b];
        }
        if (total >= n) { free(tmp); return ORC_NOT_COMPRESSIBLE; }
        for (int b = 0; b < nBlocks; ++b) {
            put32(out + 1 + 8 * b, (uint32_t)size[b]); put32(out + 5 + 8 * b, (uint32_t)res[b]);
            memcpy(out + ptr, (res[b] != size[b] ? tmp : in) + start[b], (size_t)res[b]);
            ptr += res[b];
        }
        free(tmp);
        return ptr;
    }
    for (int b = 0; b < nBlocks; ++b) {                  /* coder.cpp:111-155 */
        int room = size[b]; if (room > n - ptr) room = n - ptr;
        int r = encode_block(coder, in + start[b], out + ptr, size[b], room);
        if (r < 0) {
            if (ptr + size[b] >= n) return ORC_NOT_COMPRESSIBLE;
            r = size[b]; memcpy(out + ptr, in + start[b], (size_t)r);
        }
        put32(out + 1 + 8 * b, (uint32_t)size[b]); put32(out + 5 + 8 * b, (uint32_t)r);
        ptr += r;
    }
    return ptr;
}

int orc_coder_decompress(const unsigned char *in, unsigned char *out, int coder)
{
    if (coder < 1 || coder > 3) return ORC_BAD_PARAMETER;
    int nBlocks = in[0];
    if (nBlocks == 1) return decode_block(coder, in + 1, out);
    int inPtr = 1 + 8 * nBlocks, outPtr = 0, total = 0, err = 0;
    for (int b = 0; b < nBlocks; ++b) {
        int rawSize = (int)get32(in + 1 + 8 * b), packed = (int)get32(in + 5 + 8 * b), r;
        if (packed != rawSize) r = decode_block(coder, in + inPtr, out + outPtr);
        else { r = rawSize; memcpy(out + outPtr, in + inPtr, (size_t)rawSize); }
        if (r < 0) err = r;
        total += r; inPtr += packed; outPtr += rawSize;
    }
    return err ? err : total;
}

/* ------------------------------------------------------------------------------------------ */
/* LZP decoder (the host-side preprocessing stage; only its inverse is restated: blocks made by   */
/* the reference's default options can then be decoded).  lzp.cpp:564-674 (scalar tail loop =   */
/* the definition), container lzp.cpp:813-887.                                                  */
/* ------------------------------------------------------------------------------------------ */
static int lzp_decode_block(const unsigned char *in, const unsigned char *inEnd, unsigned char *out, long outCap, int hashSize, int minLen)
{
    if (inEnd - in < 4) return ORC_UNEXPECTED_EOB;
    int *lookup = calloc((size_t)1 << hashSize, sizeof(int));
    if (!lookup) return ORC_NOT_ENOUGH_MEMORY;
    const uint32_t mask = ((uint32_t)1 << hashSize) - 1;
    long o = 0;
    if (outCap < 4) { free(lookup); return ORC_DATA_CORRUPT; }
    for (int i = 0; i < 4; ++i) out[o++] = *in++;
    uint32_t ctx = (uint32_t)out[3] | ((uint32_t)out[2] << 8) | ((uint32_t)out[1] << 16) | ((uint32_t)out[0] << 24);
    int bad = 0;
    while (in < inEnd) {
        const uint32_t idx = ((ctx >> 15) ^ ctx ^ (ctx >> 3)) & mask;
        const int value = lookup[idx]; lookup[idx] = (int)o;
        if (*in == 0xf2 && value > 0) {                              /* LIBBSC_LZP_MATCH_FLAG */
            ++in;
            if (in >= inEnd) { bad = 1; break; }
            if (*in != 255) {
                long len = minLen;
                for (;;) { if (in >= inEnd) { bad = 1; break; } len += *in; if (*in++ != 254) break; }
                if (bad || o + len > outCap) { bad = 1; break; }
                for (long k = 0; k < len; ++k, ++o) out[o] = out[value + k];     /* forward copy: may overlap */
                ctx = (uint32_t)out[o - 1] | ((uint32_t)out[o - 2] << 8) | ((uint32_t)out[o - 3] << 16) | ((uint32_t)out[o - 4] << 24);
            } else {
                ++in;
                if (o >= outCap) { bad = 1; break; }
                out[o++] = 0xf2; ctx = (ctx << 8) | 0xf2;
            }
        } else {
            if (o >= outCap) { bad = 1; break; }
            ctx = (ctx << 8) | (out[o++] = *in++);
        }
    }
    free(lookup);
    return bad ? ORC_DATA_CORRUPT : (int)o;
}

int orc_lzp_decompress(const unsigned char *in, unsigned char *out, int n, int outCap, int hashSize, int minLen)
{
    if (n < 1) return ORC_UNEXPECTED_EOB;
    const int nBlocks = in[0];
    if (nBlocks == 1) return lzp_decode_block(in + 1, in + n, out, outCap, hashSize, minLen);
    if (nBlocks == 0 || n < 1 + 8 * nBlocks) return ORC_DATA_CORRUPT;
    long inPtr = 1 + 8L * nBlocks, outPtr = 0;
    for (int b = 0; b < nBlocks; ++b) {
        const int outSize = (int)get32(in + 1 + 8 * b), inSize = (int)get32(in + 5 + 8 * b);
        if (outSize < 0 || inSize < 0 || inPtr + inSize > n || outPtr + outSize > outCap) return ORC_DATA_CORRUPT;
        int r;
        if (inSize != outSize) r = lzp_decode_block(in + inPtr, in + inPtr + inSize, out + outPtr, outSize, hashSize, minLen);
        else { r = inSize; memcpy(out + outPtr, in + inPtr, (size_t)inSize); }
        if (r < 0) return r;
        if (r != outSize) return ORC_DATA_CORRUPT;
        inPtr += inSize; outPtr += outSize;
    }
    return (int)outPtr;
}

/* ------------------------------------------------------------------------------------------ */
/* Block framing.  libbsc.cpp:68-81, 213-338, 340-418, 522-617.  LZP is not part of the path.   */
/* ------------------------------------------------------------------------------------------ */
int orc_store(const unsigned char *in, unsigned char *out, int n)
{
    uint32_t a = orc_adler32(in, n);
    memmove(out + ORC_HEADER_SIZE, in, (size_t)n);
    put32(out, (uint32_t)(n + ORC_HEADER_SIZE)); put32(out + 4, (uint32_t)n); put32(out + 8, 0); put32(out + 12, 0);
    put32(out + 16, a); put32(out + 20, a); put32(out + 24, orc_adler32(out, 24));
    return n + ORC_HEADER_SIZE;
}

int orc_compress(const unsigned char *in, unsigned char *out, int n, int blockSorter, int coder, int features)
{
    if (!(blockSorter == 1 || (blockSorter >= 3 && blockSorter <= 8))) return ORC_BAD_PARAMETER;
    if (coder < 1 || coder > 3) return ORC_BAD_PARAMETER;
    int mode = blockSorter | (coder << 5);
    if (n < 0 || n > 1073741824) return ORC_BAD_PARAMETER;
    if (n <= ORC_HEADER_SIZE) return orc_store(in, out, n);
    memcpy(out, in, (size_t)n);

    int indexes[256]; unsigned char num_indexes = 0; int index;
    if (blockSorter == 1) index = orc_bwt_encode(out, n, &num_indexes, indexes);
    else index = orc_st_encode(out, n, blockSorter);
    if (n < 64 * 1024) num_indexes = 0;
    if (index < 0) return index;

    unsigned char *buf = malloc((size_t)n + 4096);
    if (!buf) return ORC_NOT_ENOUGH_MEMORY;
    int r = orc_coder_compress(out, buf, n, coder, features);
    if (r >= 0) memcpy(out + ORC_HEADER_SIZE, buf, (size_t)r);
    free(buf);
    if (r < 0 || r + 1 + 4 * num_indexes >= n) return orc_store(in, out, n);
    for (int t = 0; t < num_indexes; ++t) put32(out + ORC_HEADER_SIZE + r + 4 * t, (uint32_t)indexes[t]);
    out[ORC_HEADER_SIZE + r + 4 * num_indexes] = num_indexes;
    r += 1 + 4 * num_indexes;
    put32(out, (uint32_t)(r + ORC_HEADER_SIZE)); put32(out + 4, (uint32_t)n); put32(out + 8, (uint32_t)mode); put32(out + 12, (uint32_t)index);
    put32(out + 16, orc_adler32(in, n)); put32(out + 20, orc_adler32(out + ORC_HEADER_SIZE, r)); put32(out + 24, orc_adler32(out, 24));
    return r + ORC_HEADER_SIZE;
}

int orc_block_info(const unsigned char *h, int hdrSize, int *pBlockSize, int *pDataSize)
{
    if (hdrSize < ORC_HEADER_SIZE) return ORC_UNEXPECTED_EOB;
    if (get32(h + 24) != orc_adler32(h, 24)) return ORC_DATA_CORRUPT;
    int blockSize = (int)get32(h), dataSize = (int)get32(h + 4), mode = (int)get32(h + 8), index = (int)get32(h + 12);
    int lzpHash = (mode >> 16) & 0xff, lzpMin = (mode >> 8) & 0xff, coder = (mode >> 5) & 7, sorter = mode & 0x1f;
    int rebuilt = 0;
    if (sorter == 1 || (sorter >= 3 && sorter <= 8)) rebuilt = sorter; else if (sorter > 0) return ORC_DATA_CORRUPT;
    if (coder >= 1 && coder <= 3) rebuilt += coder << 5; else if (coder > 0) return ORC_DATA_CORRUPT;
    if (lzpMin != 0 || lzpHash != 0) {
        if (lzpMin < 4 || lzpHash < 10 || lzpHash > 28) return ORC_DATA_CORRUPT;
        rebuilt += (lzpMin << 8) + (lzpHash << 16);
    }
    if (rebuilt != mode) return ORC_DATA_CORRUPT;
    if (blockSize < ORC_HEADER_SIZE || blockSize > ORC_HEADER_SIZE + dataSize) return ORC_DATA_CORRUPT;
    if (index < 0 || index > dataSize) return ORC_DATA_CORRUPT;
    if (pBlockSize) *pBlockSize = blockSize;
    if (pDataSize) *pDataSize = dataSize;
    return ORC_NO_ERROR;
}

int orc_decompress(const unsigned char *in, int inSize, unsigned char *out, int outSize)
{
    int blockSize = 0, dataSize = 0;
    int info = orc_block_info(in, inSize, &blockSize, &dataSize);
    if (info != ORC_NO_ERROR) return info;
    if (inSize < blockSize || outSize < dataSize) return ORC_UNEXPECTED_EOB;
    if (get32(in + 20) != orc_adler32(in + ORC_HEADER_SIZE, blockSize - ORC_HEADER_SIZE)) return ORC_DATA_CORRUPT;
    int mode = (int)get32(in + 8);
    if (mode == 0) { memcpy(out, in + ORC_HEADER_SIZE, (size_t)dataSize); return ORC_NO_ERROR; }
    int index = (int)get32(in + 12); uint32_t adler = get32(in + 16);
    int coder = (mode >> 5) & 7, sorter = mode & 0x1f, lzpHash = (mode >> 16) & 0xff, lzpMin = (mode >> 8) & 0xff;
    int lzSize = orc_coder_decompress(in + ORC_HEADER_SIZE, out, coder);
    if (lzSize < 0) return lzSize;
    int r;
    if (sorter == 1) r = orc_bwt_decode(out, lzSize, index);
    else r = orc_st_decode(out, lzSize, sorter, index);           /* libbsc.cpp:584-589 */
    if (r < 0) return r;
    if (mode != (mode & 0xff)) {                                  /* libbsc.cpp:599-613: undo the LZP stage */
        unsigned char *lz = malloc((size_t)lzSize + 1);
        if (!lz) return ORC_NOT_ENOUGH_MEMORY;
        memcpy(lz, out, (size_t)lzSize);
        r = orc_lzp_decompress(lz, out, lzSize, dataSize, lzpHash, lzpMin);
        free(lz);
        if (r < 0) return r;
        lzSize = r;
    }
    return (lzSize == dataSize && adler == orc_adler32(out, dataSize)) ? ORC_NO_ERROR : ORC_DATA_CORRUPT;
}
