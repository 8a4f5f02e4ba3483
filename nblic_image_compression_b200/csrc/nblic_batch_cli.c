/*
 * nblic_batch_cli.c -- batch front end for libnblic_b200.so.  The reference CLI
 * (src/NBLIC_main.c) handles one image per process launch and does its file I/O inline (src/FileIO.c:81-287:
 * loadPGMfile / loadBMPfile / writePGMfile / writeBytesToFile, each a blocking fread / fwrite around the codec call).
 * This front end takes many files and runs them as a three-stage pipeline of GROUPS of files:
 *
 *      reader thread    : open + parse group g+1 straight into PINNED host staging (nblic_b200_host_alloc)
 *      main thread      : one nblic_b200_encode_batch / nblic_b200_decode_batch call on group g
 *                         (every image of the group gets its own coder stream on the GPU; copies are asynchronous
 *                         DMA because the staging is pinned)
 *      writer thread    : write the results of group g-1
 *
 * over a ring of three slots, so the wall time tends to max(read, GPU, write) instead of their sum (-v prints all
 * four).  A group closes when its inputs reach the staging budget (-b<MiB>, default 256) or 8192 files.
 * Switches follow the reference's (-c / -d, -n<N>, -e<E>, -v; src/NBLIC_main.c:52-95).
 *
 *   nblic_batch -c [-n<near>] [-e<effort>] [-g<gpu>] [-b<MiB>] [-v] <out_dir> <in.pgm|in.bmp>...   -> <out_dir>/<stem>.nblic
 *   nblic_batch -d [-g<gpu>] [-b<MiB>] [-v] <out_dir> <in.nblic>...                              -> <out_dir>/<stem>.pgm
 *
 * The file formats are read and written by this file's own small parsers (binary PGM; 8-bit palettised
 * BMP, bottom-up or top-down, rows padded to 4 bytes, the palette index taken as the gray value as the
 * reference's loader does, src/FileIO.c:170-245).  Host side only; all coding happens on the GPU.
 */
#include <ctype.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/nblic_b200.h"

#define N_SLOTS 3
#define MAX_GROUP_FILES 8192

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

/* ---- pinned arenas ------------------------------------------------------------------------------ */
typedef struct { uint8_t *p; size_t cap, used; } arena_t;

static int arena_reserve(arena_t *a, size_t bytes) { /* only ever called on an empty arena (used == 0) */
    if (bytes <= a->cap) return 0;
    if (a->p) nblic_b200_host_free(a->p);
    a->p = (uint8_t *)nblic_b200_host_alloc(bytes);
    a->cap = a->p ? bytes : 0;
    return a->p ? 0 : -1;
}
static size_t round16(size_t v) { return (v + 15) & ~(size_t)15; }

/* ---- one group of files in flight --------------------------------------------------------------- */
enum { SLOT_FREE, SLOT_LOADED, SLOT_CODED };
typedef struct {
    int state, first, count;
    arena_t in, out;
    size_t in_off[MAX_GROUP_FILES], in_len[MAX_GROUP_FILES], out_off[MAX_GROUP_FILES], out_cap[MAX_GROUP_FILES], out_len[MAX_GROUP_FILES];
    int hs[MAX_GROUP_FILES], ws[MAX_GROUP_FILES], status[MAX_GROUP_FILES];
} slot_t;

typedef struct {
    int decode, near, effort, verbose, n_files, error, failed_files, n_groups; /* n_groups: -1 until the reader has seen the last file */
    size_t budget;
    const char *dir;
    char **files;
    slot_t *slot;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    double read_s, write_s, gpu_s, pixels, bytes;
} job_t;

static void set_state(job_t *j, slot_t *s, int st) { pthread_mutex_lock(&j->mu); s->state = st; pthread_cond_broadcast(&j->cv); pthread_mutex_unlock(&j->mu); }
static void wait_state(job_t *j, slot_t *s, int st) { pthread_mutex_lock(&j->mu); while (s->state != st && !j->error) pthread_cond_wait(&j->cv, &j->mu); pthread_mutex_unlock(&j->mu); }
static void raise_error(job_t *j) { pthread_mutex_lock(&j->mu); j->error = 1; pthread_cond_broadcast(&j->cv); pthread_mutex_unlock(&j->mu); }

/* ---- file formats --------------------------------------------------------------------------------- */
static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* binary PGM header: "P5" <ws> width <ws> height <ws> maxval(<256) <single ws>; '#' comments allowed.  Returns the raster offset or 0. */
static size_t pgm_header(const uint8_t *d, size_t len, int *h, int *w) {
    size_t p = 2;
    long v[3];
    int k;
    if (len < 7 || d[0] != 'P' || d[1] != '5') return 0;
    for (k = 0; k < 3; k++) {
        for (;;) {
            while (p < len && isspace(d[p])) p++;
            if (p < len && d[p] == '#') { while (p < len && d[p] != '\n') p++; continue; }
            break;
        }
        if (p >= len || !isdigit(d[p])) return 0;
        v[k] = 0;
        while (p < len && isdigit(d[p])) { v[k] = v[k] * 10 + (d[p] - '0'); if (v[k] > 100000000L) return 0; p++; }
    }
    if (p >= len || !isspace(d[p])) return 0;
    p++;
    if (v[0] <= 0 || v[1] <= 0 || v[2] <= 0 || v[2] > 255) return 0;
    *w = (int)v[0]; *h = (int)v[1];
    return p;
}

typedef struct { int h, w, bottom_up; size_t data_off, stride; } bmp_t;
static int bmp_header(const uint8_t *d, size_t len, bmp_t *b) {
    uint32_t bpp, compression;
    int32_t bw, bh;
    if (len < 54 || d[0] != 'B' || d[1] != 'M') return -1;
    b->data_off = le32(d + 10);
    bw = (int32_t)le32(d + 18); bh = (int32_t)le32(d + 22);
    bpp = d[28] | ((uint32_t)d[29] << 8);
    compression = le32(d + 30);
    if (bpp != 8 || compression != 0 || bw <= 0 || bh == 0 || bh == INT32_MIN) return -1;
    b->w = bw; b->h = bh < 0 ? -bh : bh; b->bottom_up = bh > 0;
    b->stride = ((size_t)bw + 3) & ~(size_t)3;
    return 0;
}

static long file_size(FILE *f) {
    long n;
    if (fseek(f, 0, SEEK_END) != 0 || (n = ftell(f)) < 0 || fseek(f, 0, SEEK_SET) != 0) return -1;
    return n;
}

/* What loading `path` will need: input bytes in staging (raster for encode, file for decode), dims. */
typedef struct { const char *path; long size; int h, w, kind /* 0 pgm, 1 bmp, 2 stream */; size_t raster_off; bmp_t bmp; } probe_t;

static int probe_file(const job_t *j, const char *path, probe_t *pr) { /* header only; the file is reopened by load_payload (a group may hold thousands) */
    uint8_t head[4096];
    size_t got;
    FILE *f = fopen(path, "rb");
    memset(pr, 0, sizeof *pr);
    pr->path = path;
    if (!f || (pr->size = file_size(f)) < 0) { fprintf(stderr, "  ***Error : open %s failed\n", path); if (f) fclose(f); return -1; }
    got = fread(head, 1, sizeof head, f);
    fclose(f);
    if (j->decode) {
        int n_ = 0, e_ = 0;
        pr->kind = 2;
        if (nblic_b200_peek(head, got, &pr->h, &pr->w, &n_, &e_) != 0) { fprintf(stderr, "  ***Error : %s is not a .nblic stream\n", path); return -1; }
        return 0;
    }
    pr->raster_off = pgm_header(head, got, &pr->h, &pr->w);
    if (pr->raster_off && (size_t)pr->h * (size_t)pr->w <= (size_t)pr->size - pr->raster_off) { pr->kind = 0; return 0; }
    if (bmp_header(head, got, &pr->bmp) == 0 && pr->bmp.data_off <= (size_t)pr->size &&
        pr->bmp.stride * (size_t)pr->bmp.h <= (size_t)pr->size - pr->bmp.data_off) {
        pr->kind = 1; pr->h = pr->bmp.h; pr->w = pr->bmp.w;
        return 0;
    }
    fprintf(stderr, "  ***Error : %s is neither a binary PGM nor an 8-bit BMP\n", path);
    return -1;
}

/* Read the payload of a probed file into dst (pinned staging). */
static int load_payload(const probe_t *pr, uint8_t *dst) {
    FILE *f = fopen(pr->path, "rb");
    int ok = f != NULL;
    if (!ok) return -1;
    if (pr->kind == 2) ok = fread(dst, 1, (size_t)pr->size, f) == (size_t)pr->size;
    else if (pr->kind == 0) ok = fseek(f, (long)pr->raster_off, SEEK_SET) == 0 && fread(dst, 1, (size_t)pr->h * pr->w, f) == (size_t)pr->h * pr->w;
    else {
        int r; /* file rows in file order (one forward pass); file row r is image row h-1-r when bottom-up */
        const size_t padding = pr->bmp.stride - (size_t)pr->w;
        uint8_t pad[4];
        ok = fseek(f, (long)pr->bmp.data_off, SEEK_SET) == 0;
        for (r = 0; r < pr->h && ok; r++) {
            const int y = pr->bmp.bottom_up ? pr->h - 1 - r : r;
            ok = fread(dst + (size_t)y * pr->w, 1, (size_t)pr->w, f) == (size_t)pr->w && (padding == 0 || r == pr->h - 1 || fread(pad, 1, padding, f) == padding);
        }
    }
    fclose(f);
    return ok ? 0 : -1;
}

static void out_path(char *dst, size_t cap, const char *dir, const char *in, const char *suffix) {
    const char *base = strrchr(in, '/');
    const char *dot;
    size_t stem;
    base = base ? base + 1 : in;
    dot = strrchr(base, '.');
    stem = dot ? (size_t)(dot - base) : strlen(base);
    snprintf(dst, cap, "%s/%.*s%s", dir, (int)stem, base, suffix);
}

/* ---- stage 1: reader ------------------------------------------------------------------------------ */
static void *reader_main(void *arg) {
    job_t *j = (job_t *)arg;
    int next = 0, g = 0;
    probe_t *probes = (probe_t *)calloc(MAX_GROUP_FILES, sizeof *probes);
    if (!probes) { raise_error(j); return NULL; }
    while (next < j->n_files && !j->error) {
        slot_t *s = &j->slot[g % N_SLOTS];
        size_t in_total = 0, out_total = 0;
        int k, cnt = 0;
        double t0;
        wait_state(j, s, SLOT_FREE);
        if (j->error) break;
        t0 = now_s();
        /* size the group: probe headers until the budget is reached (the first file always fits) */
        while (next + cnt < j->n_files && cnt < MAX_GROUP_FILES) {
            probe_t *pr = &probes[cnt];
            size_t in_b, out_b;
            if (probe_file(j, j->files[next + cnt], pr) != 0) { raise_error(j); break; }
            in_b = j->decode ? (size_t)pr->size : (size_t)pr->h * pr->w;
            out_b = j->decode ? (size_t)pr->h * pr->w : nblic_b200_stream_bound(pr->h, pr->w);
            if (cnt > 0 && in_total + in_b > j->budget) break;
            s->in_off[cnt] = in_total; s->in_len[cnt] = in_b; s->out_off[cnt] = out_total; s->out_cap[cnt] = out_b;
            s->hs[cnt] = pr->h; s->ws[cnt] = pr->w;
            in_total += round16(in_b + 16); out_total += round16(out_b);
            cnt++;
        }
        if (j->error) break;
        if (arena_reserve(&s->in, in_total) != 0 || arena_reserve(&s->out, out_total) != 0) {
            fprintf(stderr, "  ***Error : pinned staging of %zu + %zu bytes failed\n", in_total, out_total);
            raise_error(j);
            break;
        }
        for (k = 0; k < cnt; k++) {
            if (load_payload(&probes[k], s->in.p + s->in_off[k]) != 0) { fprintf(stderr, "  ***Error : read %s failed\n", j->files[next + k]); raise_error(j); break; }
            memset(s->in.p + s->in_off[k] + s->in_len[k], 0, 16);
        }
        if (j->error) break;
        s->first = next; s->count = cnt;
        next += cnt;
        g++;
        pthread_mutex_lock(&j->mu);
        j->read_s += now_s() - t0;
        if (next >= j->n_files) j->n_groups = g;
        s->state = SLOT_LOADED;
        pthread_cond_broadcast(&j->cv);
        pthread_mutex_unlock(&j->mu);
    }
    free(probes);
    return NULL;
}

/* ---- stage 3: writer ------------------------------------------------------------------------------ */
static int write_file(const char *path, const void *head, size_t head_len, const void *body, size_t body_len) {
    FILE *f = fopen(path, "wb");
    int ok;
    if (!f) return -1;
    ok = (head_len == 0 || fwrite(head, 1, head_len, f) == head_len) && fwrite(body, 1, body_len, f) == body_len;
    return fclose(f) == 0 && ok ? 0 : -1;
}

static void *writer_main(void *arg) {
    job_t *j = (job_t *)arg;
    int g;
    for (g = 0;; g++) {
        slot_t *s = &j->slot[g % N_SLOTS];
        int k, done, failed = 0;
        double t0, px = 0, by = 0;
        pthread_mutex_lock(&j->mu);
        while (s->state != SLOT_CODED && !j->error && !(j->n_groups >= 0 && g >= j->n_groups)) pthread_cond_wait(&j->cv, &j->mu);
        done = j->error || (j->n_groups >= 0 && g >= j->n_groups);
        pthread_mutex_unlock(&j->mu);
        if (done) break;
        t0 = now_s();
        for (k = 0; k < s->count; k++) {
            char path[4096], head[64];
            const char *in = j->files[s->first + k];
            if (s->status[k] != NBLIC_B200_OK) { fprintf(stderr, "  ***Error : %s failed (status %d)\n", in, s->status[k]); failed++; continue; }
            if (!j->decode) {
                out_path(path, sizeof path, j->dir, in, ".nblic");
                if (write_file(path, NULL, 0, s->out.p + s->out_off[k], s->out_len[k]) != 0) { fprintf(stderr, "  ***Error : write %s failed\n", path); failed++; }
                by += (double)s->out_len[k];
            } else {
                const int hl = snprintf(head, sizeof head, "P5\n%d %d\n255\n", s->ws[k], s->hs[k]);
                out_path(path, sizeof path, j->dir, in, ".pgm");
                if (write_file(path, head, (size_t)hl, s->out.p + s->out_off[k], (size_t)s->hs[k] * (size_t)s->ws[k]) != 0) { fprintf(stderr, "  ***Error : write %s failed\n", path); failed++; }
                by += (double)s->in_len[k];
            }
            px += (double)s->hs[k] * s->ws[k];
            if (j->verbose) printf("  %s -> %s  %d x %d\n", in, path, s->ws[k], s->hs[k]);
        }
        pthread_mutex_lock(&j->mu);
        j->write_s += now_s() - t0;
        j->pixels += px; j->bytes += by; j->failed_files += failed;
        s->state = SLOT_FREE;
        pthread_cond_broadcast(&j->cv);
        pthread_mutex_unlock(&j->mu);
    }
    return NULL;
}

/* ---- stage 2 (main thread): the batch calls ---------------------------------------------------------- */
int main(int argc, char **argv) {
    static job_t job; /* zero-initialised */
    job_t *j = &job;
    int gpu = 0, a = 1, g, k, mib = 256;
    nblic_b200_ctx *ctx;
    pthread_t reader, writer;
    const uint8_t **in_ptr;
    uint8_t **out_ptr;
    int *nears, *efforts;
    double t0, wall;

    j->decode = -1; j->effort = 1; j->n_groups = -1;
    for (; a < argc && argv[a][0] == '-'; a++) { /* compact switches, any order, as the reference CLI */
        const char *s = argv[a] + 1;
        for (; *s; s++) {
            if (*s == 'c') j->decode = 0;
            else if (*s == 'd') j->decode = 1;
            else if (*s == 'v' || *s == 'V') j->verbose = 1;
            else if (*s == 't') { /* accepted and ignored: every image already gets its own coder stream */ }
            else if (*s == 'n' || *s == 'e' || *s == 'g' || *s == 'b') {
                int val = 0, digits = 0;
                const char which = *s;
                while (isdigit((unsigned char)s[1])) { val = val * 10 + (s[1] - '0'); s++; digits++; }
                if (!digits) { fprintf(stderr, "-%c needs a number\n", which); return -1; }
                if (which == 'n') j->near = val; else if (which == 'e') j->effort = val; else if (which == 'g') gpu = val; else mib = val;
            } else { fprintf(stderr, "unknown switch -%c\n", *s); return -1; }
        }
    }
    if (j->decode < 0 || argc - a < 2) {
        fprintf(stderr, "usage: %s -c [-n<near>] [-e<effort>] [-g<gpu>] [-b<MiB per group>] [-v] <out_dir> <in.pgm|in.bmp>...\n"
                        "       %s -d [-g<gpu>] [-b<MiB per group>] [-v] <out_dir> <in.nblic>...\n", argv[0], argv[0]);
        return -1;
    }
    j->dir = argv[a++];
    j->files = argv + a;
    j->n_files = argc - a;
    j->budget = mib > 0 ? (size_t)mib << 20 : 1; /* -b0: one file per group (exercises the pipeline on small inputs) */
    j->slot = (slot_t *)calloc(N_SLOTS, sizeof(slot_t));
    in_ptr = (const uint8_t **)calloc(MAX_GROUP_FILES, sizeof *in_ptr); out_ptr = (uint8_t **)calloc(MAX_GROUP_FILES, sizeof *out_ptr);
    nears = (int *)calloc(MAX_GROUP_FILES, sizeof(int)); efforts = (int *)calloc(MAX_GROUP_FILES, sizeof(int));
    if (!j->slot || !in_ptr || !out_ptr || !nears || !efforts) return -1;
    pthread_mutex_init(&j->mu, NULL);
    pthread_cond_init(&j->cv, NULL);

    ctx = nblic_b200_create(gpu); /* before the threads: pinned allocations need the CUDA context */
    if (!ctx) { fprintf(stderr, "  ***Error : %s\n", nblic_b200_last_error(NULL)); return -1; }
    t0 = now_s();
    if (pthread_create(&reader, NULL, reader_main, j) != 0 || pthread_create(&writer, NULL, writer_main, j) != 0) return -1;
    for (g = 0;; g++) {
        slot_t *s = &j->slot[g % N_SLOTS];
        int done, failed;
        double t1;
        pthread_mutex_lock(&j->mu);
        while (s->state != SLOT_LOADED && !j->error && !(j->n_groups >= 0 && g >= j->n_groups)) pthread_cond_wait(&j->cv, &j->mu);
        done = j->error || (j->n_groups >= 0 && g >= j->n_groups);
        pthread_mutex_unlock(&j->mu);
        if (done) break;
        for (k = 0; k < s->count; k++) { in_ptr[k] = s->in.p + s->in_off[k]; out_ptr[k] = s->out.p + s->out_off[k]; }
        t1 = now_s();
        if (!j->decode)
            failed = nblic_b200_encode_batch(ctx, s->count, in_ptr, s->hs, s->ws, j->near, j->effort, out_ptr, s->out_cap, s->out_len, NULL, s->status);
        else
            failed = nblic_b200_decode_batch(ctx, s->count, in_ptr, s->in_len, out_ptr, s->out_cap, s->hs, s->ws, nears, efforts, s->status);
        j->gpu_s += now_s() - t1;
        if (failed < 0) { fprintf(stderr, "  ***Error : %s\n", nblic_b200_last_error(ctx)); raise_error(j); break; }
        set_state(j, s, SLOT_CODED);
    }
    pthread_join(reader, NULL);
    pthread_join(writer, NULL);
    wall = now_s() - t0;
    if (j->verbose)
        printf("  %d images in %d groups, %.3f MPixel, %.0f stream bytes (%.4f bpp), %d failed\n"
               "  wall %.3f s (%.1f MPixel/s) = pipeline of read %.3f s | GPU batch calls %.3f s | write %.3f s\n",
               j->n_files, j->n_groups, j->pixels / 1e6, j->bytes, j->pixels > 0 ? 8.0 * j->bytes / j->pixels : 0.0, j->failed_files,
               wall, j->pixels / 1e6 / (wall > 0 ? wall : 1), j->read_s, j->gpu_s, j->write_s);
    for (k = 0; k < N_SLOTS; k++) { if (j->slot[k].in.p) nblic_b200_host_free(j->slot[k].in.p); if (j->slot[k].out.p) nblic_b200_host_free(j->slot[k].out.p); }
    nblic_b200_destroy(ctx);
    return j->error || j->failed_files ? -1 : 0;
}
