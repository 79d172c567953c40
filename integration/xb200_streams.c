/* xb200_streams.c -- N independent streams through the xeve C API (inc/xeve.h) in one process, one host thread per stream.
 *
 * An application-level program: it uses nothing but the public API of the library it is linked with (the drop-in
 * libxeve_b200_dropin.so, or the unmodified reference library -- the code is the same), the way app/xeve_app.c drives one stream:
 * xeve_create, then xeve_push / xeve_encode per frame, XEVE_CFG_SET_FORCE_OUT at the end, bitstream written as it comes out
 * (app/xeve_app.c:1130-1290).  With the drop-in library every stream's pictures are decided on the B200 and the streams share the
 * device; the host threads only push frames and entropy-code.  bench.py times it.
 *
 *   xb200_streams -i clip.yuv -w 1920 -h 1080 -d 8 -z 33 -n 4 -m 8 --preset fast [-q 32] [-o /dev/shm/out] [-x "name=value;.."] [-r passes]
 *
 * Prints one JSON line: wall seconds of the whole job (first push to last byte), per stream the seconds inside xeve_encode and the
 * bitstream size; stream k's bitstream goes to <out>.<k>.evc when -o is given. */
#define _GNU_SOURCE
#include "xeve.h"
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef XS_WITH_STATS
#include "xeve_b200_engine.h"
#endif

typedef struct {
    int            k, w, h, depth, frames, threads, preset, qp;
    const char    *extra, *out;
    const uint8_t *yuv;
    pthread_barrier_t *start;
    double         enc_s, push_s, total_s;
    int64_t        bytes;
    int            err;
    double         chain_ms, wait_ms;
    int64_t        n_cu, device_pictures;
} Stream;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static int img_addref(XEVE_IMGB *i) { return ++i->refcnt; }
static int img_getref(XEVE_IMGB *i) { return i->refcnt; }
static int img_release(XEVE_IMGB *i) { return --i->refcnt; }

static void *run_stream(void *arg)
{
    Stream   *s = (Stream *)arg;
    XEVE_CDSC cdsc;
    memset(&cdsc, 0, sizeof(cdsc));
    XEVE_PARAM *p = &cdsc.param;
    xeve_param_default(p);
    xeve_param_ppt(p, XEVE_PROFILE_BASELINE, s->preset, XEVE_TUNE_NONE);
    p->w = s->w; p->h = s->h; p->fps.num = 30; p->fps.den = 1;
    p->threads = s->threads;
    if(s->qp >= 0) p->qp = s->qp;
    p->cs = XEVE_CS_SET(XEVE_CF_YCBCR420, p->codec_bit_depth, 0);
    if(s->extra && *s->extra) {
        char *dup = strdup(s->extra), *save = NULL;
        for(char *tok = strtok_r(dup, ";", &save); tok; tok = strtok_r(NULL, ";", &save)) {
            char *eq = strchr(tok, '=');
            if(!eq) continue;
            *eq = 0;
            if(xeve_param_parse(p, tok, eq + 1) != XEVE_OK) fprintf(stderr, "xb200_streams: bad parameter %s\n", tok);
        }
        free(dup);
    }
    cdsc.max_bs_buf_size = 32 * 1024 * 1024;
    int  err = 0;
    XEVE id = xeve_param_check(p) == XEVE_OK ? xeve_create(&cdsc, &err) : NULL;
    pthread_barrier_wait(s->start);          /* every stream starts pushing at the same time */
    if(!id) { s->err = err ? err : -1; return NULL; }
    FILE *fo = NULL;
    if(s->out) {
        char name[1024];
        snprintf(name, sizeof(name), "%s.%d.evc", s->out, s->k);
        fo = fopen(name, "wb");
    }
    const int    bps = s->depth > 8 ? 2 : 1;
    const size_t fsz = (size_t)s->w * s->h * 3 / 2 * bps;
    uint8_t     *bs = malloc(32 * 1024 * 1024);
    XEVE_BITB    bitb;
    XEVE_STAT    stat;
    XEVE_IMGB    img;
    memset(&bitb, 0, sizeof(bitb));
    bitb.addr = bs; bitb.bsize = 32 * 1024 * 1024;
    int          pushed = 0, bumping = 0, ret;
    const double t_begin = now_s();
    for(;;) {
        if(!bumping) {
            if(pushed < s->frames) {
                const uint8_t *f = s->yuv + fsz * pushed;
                memset(&img, 0, sizeof(img));
                img.cs = XEVE_CS_SET(XEVE_CF_YCBCR420, s->depth, 0);
                img.np = 3;
                for(int c = 0; c < 3; c++) {
                    const int cw = c ? s->w / 2 : s->w, ch = c ? s->h / 2 : s->h;
                    img.w[c] = img.aw[c] = cw; img.h[c] = img.ah[c] = ch; img.s[c] = cw * bps; img.e[c] = ch;
                }
                img.a[0] = (void *)f; img.a[1] = (void *)(f + (size_t)s->w * s->h * bps); img.a[2] = (void *)(f + (size_t)s->w * s->h * bps * 5 / 4);
                img.addref = img_addref; img.getref = img_getref; img.release = img_release; img.refcnt = 1;
                img.ts[XEVE_TS_PTS] = pushed;
                const double t0 = now_s();
                ret = xeve_push(id, &img);
                s->push_s += now_s() - t0;
                if(XEVE_FAILED(ret)) { s->err = ret; break; }
                pushed++;
            }
            else {
                int val = 1, size = sizeof(int);
                xeve_config(id, XEVE_CFG_SET_FORCE_OUT, &val, &size);
                bumping = 1;
            }
        }
        const double t0 = now_s();
        ret = xeve_encode(id, &bitb, &stat);
        s->enc_s += now_s() - t0;
        if(XEVE_FAILED(ret)) { s->err = ret; break; }
        if(ret == XEVE_OK_NO_MORE_FRM) break;
        if(ret == XEVE_OK && stat.write > 0) {
            if(fo) fwrite(bs, 1, (size_t)stat.write, fo);
            s->bytes += stat.write;
        }
    }
    s->total_s = now_s() - t_begin;
#ifdef XS_WITH_STATS
    {
        xeve_b200_stats st;
        if(xeve_b200_get_stats(id, &st) == 0) {
            s->chain_ms = st.chain_ms; s->wait_ms = st.wait_ms; s->n_cu = st.n_inter + st.n_intra; s->device_pictures = st.device_path ? st.pictures : 0;
        }
    }
#endif
    if(fo) fclose(fo);
    free(bs);
    xeve_delete(id);
    return NULL;
}

int main(int argc, char **argv)
{
    const char *in = NULL, *out = NULL, *extra = "", *preset = "fast";
    int w = 0, h = 0, depth = 8, frames = 0, n = 1, threads = 1, qp = -1, repeats = 1;
    for(int i = 1; i < argc; i++) {
        const char *a = argv[i], *v = i + 1 < argc ? argv[i + 1] : NULL;
        if(!v) { fprintf(stderr, "xb200_streams: %s needs a value\n", a); return 2; }
        if(!strcmp(a, "-i")) in = v;
        else if(!strcmp(a, "-o")) out = v;
        else if(!strcmp(a, "-w")) w = atoi(v);
        else if(!strcmp(a, "-h")) h = atoi(v);
        else if(!strcmp(a, "-d")) depth = atoi(v);
        else if(!strcmp(a, "-z")) frames = atoi(v);
        else if(!strcmp(a, "-n")) n = atoi(v);
        else if(!strcmp(a, "-m")) threads = atoi(v);
        else if(!strcmp(a, "-q")) qp = atoi(v);
        else if(!strcmp(a, "-x")) extra = v;
        else if(!strcmp(a, "-r")) repeats = atoi(v);
        else if(!strcmp(a, "--preset")) preset = v;
        else { fprintf(stderr, "xb200_streams: unknown option %s\n", a); return 2; }
        i++;
    }
    if(!in || w <= 0 || h <= 0 || frames <= 0 || n <= 0 || n > 64) { fprintf(stderr, "usage: xb200_streams -i clip.yuv -w W -h H [-d 8|10] -z frames -n streams -m threads [--preset fast|medium] [-q qp] [-o prefix] [-x name=value;..]\n"); return 2; }
    const int    ps = !strcmp(preset, "fast") ? XEVE_PRESET_FAST : !strcmp(preset, "medium") ? XEVE_PRESET_MEDIUM : !strcmp(preset, "slow") ? XEVE_PRESET_SLOW : XEVE_PRESET_PLACEBO;
    const size_t fsz = (size_t)w * h * 3 / 2 * (depth > 8 ? 2 : 1);
    uint8_t     *yuv = malloc(fsz * frames);
    FILE        *fi = fopen(in, "rb");
    if(!fi || !yuv || fread(yuv, fsz, (size_t)frames, fi) != (size_t)frames) { fprintf(stderr, "xb200_streams: cannot read %d frames from %s\n", frames, in); return 1; }
    fclose(fi);
    int bad = 0;
    for(int rep = 0; rep < repeats; rep++) {   /* the whole job again (encoders created anew): one JSON line per pass */
        Stream           *st = calloc((size_t)n, sizeof(Stream));
        pthread_t        *th = calloc((size_t)n, sizeof(pthread_t));
        pthread_barrier_t start;
        pthread_barrier_init(&start, NULL, (unsigned)n + 1);
        for(int k = 0; k < n; k++) {
            Stream *s = &st[k];
            s->k = k; s->w = w; s->h = h; s->depth = depth; s->frames = frames; s->threads = threads; s->preset = ps; s->qp = qp;
            s->extra = extra; s->out = out; s->yuv = yuv; s->start = &start;
            pthread_create(&th[k], NULL, run_stream, s);
        }
        pthread_barrier_wait(&start);           /* all encoders created (device contexts, buffers): the job starts here */
        const double t0 = now_s();
        for(int k = 0; k < n; k++) pthread_join(th[k], NULL);
        const double wall = now_s() - t0;
        printf("{\"streams\": %d, \"frames\": %d, \"threads\": %d, \"wall_s\": %.6f, \"pictures_per_s\": %.4f, \"per_stream\": [", n, frames, threads, wall,
               (double)n * frames / wall);
        for(int k = 0; k < n; k++) {
            const Stream *s = &st[k];
            bad |= s->err != 0;
            printf("%s{\"err\": %d, \"enc_s\": %.6f, \"push_s\": %.6f, \"total_s\": %.6f, \"bytes\": %lld, \"chain_ms\": %.3f, \"wait_ms\": %.3f, \"cu_analyses\": %lld, "
                   "\"device_pictures\": %lld}", k ? ", " : "", s->err, s->enc_s, s->push_s, s->total_s, (long long)s->bytes, s->chain_ms, s->wait_ms,
                   (long long)s->n_cu, (long long)s->device_pictures);
        }
        printf("]}\n");
        fflush(stdout);
        pthread_barrier_destroy(&start);
        free(st); free(th);
    }
    return bad ? 1 : 0;
}
