/* g2gen.c -- seeded synthetic corpus generators for the BASELINE.json configs.
 *
 * Generator "G2" (SURVEY.md 8(d), cfg2): enwik-shape text.
 *   vocabulary V = 50 000 pseudo-words, word length max(1,int(N(5.5,2.5))), letters drawn
 *   with English letter frequencies; tokens sampled Zipf(s=1) over V; per token 10 % replaced
 *   by a fresh random word (N(7,3) letters), 3 % by a decimal number of 1-5 digits, 2 %
 *   capitalised; tokens joined by one space, ".\n" every 100 000 tokens.
 *
 * The survey's prose spec names Python's random.Random; 4 GiB cannot be produced that way in
 * bench time, so this is the same distributional spec on a counter-based integer PRNG
 * (splitmix64), generated in independent 1 MiB chunks (chunk c of stream `seed` depends only on
 * (seed, c)) so that any byte range can be produced in parallel, on any rank, identically.
 * All arithmetic is integer (normal deviates are 12-uniform Irwin-Hall sums) => bit-identical
 * output on every host.
 *
 * Also: the cfg5 mixed-entropy classes (text / 8-bit noise / 16-symbol skewed / tiled).
 *
 * Build: gcc -O2 -shared -fPIC -pthread -o libb200z_corpus.so g2gen.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define G2_VOCAB 50000
#define G2_CHUNK (1u << 20)

static inline uint64_t sm64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t rnd_below(uint64_t *s, uint32_t n) { return (uint32_t)(((sm64(s) >> 32) * (uint64_t)n) >> 32); }

/* N(mu, sigma) in 1/1024 fixed point via Irwin-Hall(12): sum of 12 U[0,1) - 6 ~ N(0,1). */
static inline int normal_fx(uint64_t *s, int mu_fx, int sigma_fx) {
    int64_t acc = 0;
    for (int i = 0; i < 6; i++) { uint64_t r = sm64(s); acc += (int64_t)(r & 0xFFFFFFFFu) + (int64_t)(r >> 32); }
    /* acc in [0, 12*2^32); z = acc/2^32 - 6 */
    int64_t z_fx = (acc >> 22) - 6 * 1024;               /* z * 1024 */
    return mu_fx + (int)((z_fx * sigma_fx) >> 10);
}

/* English letter frequencies (per 100 000), a..z */
static const uint32_t kLetterFreq[26] = {
    8167, 1492, 2782, 4253, 12702, 2228, 2015, 6094, 6966, 153, 772, 4025, 2406,
    6749, 7507, 1929, 95, 5987, 6327, 9056, 2758, 978, 2360, 150, 1974, 74 };
static uint32_t g_letter_cdf[26];
static uint32_t g_letter_total;

typedef struct {
    uint64_t seed;
    uint32_t word_off[G2_VOCAB + 1];
    uint8_t *words;
    uint32_t zipf_cdf[G2_VOCAB];   /* cumulative of 2^32 * (1/k)/H_V, monotone */
} g2_vocab_t;

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static g2_vocab_t *g_vocab = NULL;

static inline uint8_t rnd_letter(uint64_t *s) {
    uint32_t r = rnd_below(s, g_letter_total);
    int lo = 0, hi = 25;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (g_letter_cdf[mid] > r) hi = mid; else lo = mid + 1; }
    return (uint8_t)('a' + lo);
}

static g2_vocab_t *g2_get_vocab(uint64_t seed) {
    pthread_mutex_lock(&g_lock);
    if (g_vocab && g_vocab->seed == seed) { pthread_mutex_unlock(&g_lock); return g_vocab; }
    if (!g_letter_total) {
        uint32_t c = 0;
        for (int i = 0; i < 26; i++) { c += kLetterFreq[i]; g_letter_cdf[i] = c; }
        g_letter_total = c;
    }
    g2_vocab_t *v = g_vocab ? g_vocab : (g2_vocab_t *)calloc(1, sizeof(*v));
    if (!v->words) v->words = (uint8_t *)malloc((size_t)G2_VOCAB * 24);
    v->seed = seed;
    uint64_t s = seed ^ 0xC0FFEE1234ull;
    uint32_t off = 0;
    for (int w = 0; w < G2_VOCAB; w++) {
        int len = normal_fx(&s, 5632, 2560) >> 10;        /* N(5.5, 2.5) */
        if (len < 1) len = 1;
        if (len > 23) len = 23;
        v->word_off[w] = off;
        for (int i = 0; i < len; i++) v->words[off++] = rnd_letter(&s);
    }
    v->word_off[G2_VOCAB] = off;
    /* Zipf(1) CDF in 32-bit fixed point, exact integer arithmetic:
       weight_k = floor(2^40 / k); cdf scaled to 2^32 by total. */
    uint64_t total = 0;
    for (int k = 1; k <= G2_VOCAB; k++) total += ((uint64_t)1 << 40) / (uint64_t)k;
    uint64_t acc = 0;
    for (int k = 1; k <= G2_VOCAB; k++) {
        acc += ((uint64_t)1 << 40) / (uint64_t)k;
        unsigned __int128 q = ((unsigned __int128)acc << 32) / total;
        v->zipf_cdf[k - 1] = (k == G2_VOCAB) ? 0xFFFFFFFFu : (uint32_t)(q > 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint64_t)q);
    }
    g_vocab = v;
    pthread_mutex_unlock(&g_lock);
    return v;
}

static inline uint32_t zipf_sample(const g2_vocab_t *v, uint64_t *s) {
    uint32_t r = (uint32_t)(sm64(s) >> 32);
    int lo = 0, hi = G2_VOCAB - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (v->zipf_cdf[mid] >= r) hi = mid; else lo = mid + 1; }
    return (uint32_t)lo;
}

/* Fill one chunk (index `chunk`) of stream `seed`; writes exactly n <= G2_CHUNK bytes. */
static void g2_chunk(const g2_vocab_t *v, uint64_t seed, uint64_t chunk, uint8_t *dst, size_t n) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + chunk * 0xD1B54A32D192ED03ull + 0x1234567ull;
    (void)sm64(&s);
    uint8_t tok[32];
    size_t o = 0;
    uint32_t ntok = 0;
    while (o < n) {
        uint32_t kind = rnd_below(&s, 100);
        int len;
        if (kind < 10) {                                   /* fresh random word */
            len = normal_fx(&s, 7168, 3072) >> 10;         /* N(7,3) */
            if (len < 1) len = 1;
            if (len > 30) len = 30;
            for (int i = 0; i < len; i++) tok[i] = rnd_letter(&s);
        } else if (kind < 13) {                            /* decimal number, 1-5 digits */
            len = 1 + (int)rnd_below(&s, 5);
            for (int i = 0; i < len; i++) tok[i] = (uint8_t)('0' + rnd_below(&s, 10));
        } else {
            uint32_t w = zipf_sample(v, &s);
            len = (int)(v->word_off[w + 1] - v->word_off[w]);
            memcpy(tok, v->words + v->word_off[w], (size_t)len);
            if (kind < 15) tok[0] = (uint8_t)(tok[0] - 'a' + 'A');   /* 2 % capitalised */
        }
        for (int i = 0; i < len && o < n; i++) dst[o++] = tok[i];
        ntok++;
        if (ntok % 100000u == 0) { if (o < n) dst[o++] = '.'; if (o < n) dst[o++] = '\n'; }
        else if (o < n) dst[o++] = ' ';
    }
}

typedef struct { const g2_vocab_t *v; uint64_t seed; uint64_t first_chunk; uint8_t *dst; size_t total; int tid, nthr; } job_t;

static void *g2_worker(void *arg) {
    job_t *j = (job_t *)arg;
    size_t nchunks = (j->total + G2_CHUNK - 1) / G2_CHUNK;
    for (size_t c = (size_t)j->tid; c < nchunks; c += (size_t)j->nthr) {
        size_t off = c * (size_t)G2_CHUNK;
        size_t n = j->total - off < G2_CHUNK ? j->total - off : G2_CHUNK;
        g2_chunk(j->v, j->seed, j->first_chunk + c, j->dst + off, n);
    }
    return NULL;
}

/* Generate `nbytes` of G2 text of stream `seed` starting at byte offset `offset`
 * (offset must be a multiple of 1 MiB).  Returns 0 on success. */
int b200z_corpus_g2(uint64_t seed, uint64_t offset, uint8_t *dst, uint64_t nbytes, int nthreads) {
    if (offset % G2_CHUNK) return -1;
    const g2_vocab_t *v = g2_get_vocab(seed);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; job_t jobs[256];
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (job_t){ v, seed, offset / G2_CHUNK, dst, (size_t)nbytes, t, nthreads };
        if (t) pthread_create(&th[t], NULL, g2_worker, &jobs[t]);
    }
    g2_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

/* cfg5 "mixed-entropy" file classes; cls = 0 text, 1 noise (8-bit LCG),
 * 2 = 16-symbol skewed bytes, 3 = repeated 4 KiB tile with 1 % noise. */
int b200z_corpus_class(uint64_t seed, int cls, uint8_t *dst, uint64_t nbytes) {
    uint64_t s = seed * 0xA24BAED4963EE407ull + (uint64_t)cls;
    (void)sm64(&s);
    if (cls == 0) {
        const g2_vocab_t *v = g2_get_vocab(20260922ull);
        uint64_t done = 0;
        while (done < nbytes) {
            size_t n = nbytes - done < G2_CHUNK ? (size_t)(nbytes - done) : G2_CHUNK;
            g2_chunk(v, seed, done / G2_CHUNK + 7777, dst + done, n);
            done += n;
        }
    } else if (cls == 1) {
        for (uint64_t i = 0; i < nbytes; i++) dst[i] = (uint8_t)(sm64(&s) >> 56);
    } else if (cls == 2) {
        /* 16 symbols, geometric-ish skew: symbol k with weight 2^-(k/2) */
        static const uint16_t cdf[16] = { 9362, 15982, 20663, 23973, 26314, 27969, 29139, 29967,
                                          30552, 30966, 31259, 31466, 31612, 31715, 31788, 32768 };
        for (uint64_t i = 0; i < nbytes; i++) {
            uint32_t r = (uint32_t)(sm64(&s) >> 49);      /* 15 bits */
            int k = 0; while (cdf[k] <= r) k++;
            dst[i] = (uint8_t)(0x40 + 3 * k);
        }
    } else {
        uint8_t tile[4096];
        for (int i = 0; i < 4096; i++) tile[i] = (uint8_t)(sm64(&s) >> 56);
        for (uint64_t i = 0; i < nbytes; i++) {
            uint8_t b = tile[i & 4095];
            if (rnd_below(&s, 100) == 0) b = (uint8_t)(sm64(&s) >> 56);
            dst[i] = b;
        }
    }
    return 0;
}
