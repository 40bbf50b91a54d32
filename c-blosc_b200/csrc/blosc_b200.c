/*
 * blosc_b200.c -- host side of libblosc_b200: the c-blosc 1.x C API and the chunk
 * framing / planning, in plain C, over the device layer of b2_backend.h.
 *
 * What stays on the host (it is the format contract, a few dozen integer operations per
 * call): argument validation and return codes (reference blosc/blosc.c:1062-1145,
 * 1435-1518), compute_blocksize (:962-1060), split_block (:929-959), the 16-byte header
 * (:1148-1247) and the MEMCPYED decisions (:1219-1229, :1264-1272).  Everything that
 * touches the payload -- filters, codecs, the block scheduler and the compaction of
 * variable-size blocks -- runs as CUDA kernels (dev_*.cuh); there is no CPU codec here.
 */
#include "../../include/blosc_b200.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b2_backend.h"

#define MIN_BUFFERSIZE 128        /* blosc.c:73 */
#define MAX_SPLITS 16             /* blosc.c:76 */
#define L1_SIZE (32 * 1024)       /* blosc.c:79 */

/* ---- process-global state of the non-ctx API (blosc.c:143-150) ---- */
static int g_compressor = BLOSC_BLOSCLZ;
static int g_threads = 1;
static int g_force_blocksize = 0;
static int g_initlib = 0;
static int g_splitmode = BLOSC_FORWARD_COMPAT_SPLIT;
static pthread_mutex_t g_global_mutex = PTHREAD_MUTEX_INITIALIZER;

/* ---- little-endian header accessors (blosc.c:243-289) ---- */
static int32_t rd_i32(const uint8_t* p) {
  return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}
static void wr_i32(uint8_t* p, int32_t v) {
  uint32_t u = (uint32_t)v;
  p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)(u >> 16); p[3] = (uint8_t)(u >> 24);
}

/* ------------------------------------------------------------------------- */
/* pageable-host staging: a small pool of copy threads + pinned bounce slices */
/* ------------------------------------------------------------------------- */
/* Callers of the C API usually hand in malloc'ed (pageable) buffers.  cudaMemcpy from pageable
 * memory runs at a fraction of the PCIe rate, so such buffers go through page-locked bounce
 * slices instead: a few host threads copy slice i+1 while the DMA engine moves slice i. */
#define B2_STAGE_SLICE ((size_t)8 << 20)
#define B2_STAGE_DEPTH 4
#define B2_COPY_THREADS_MAX 16

typedef struct {
  pthread_mutex_t mu;
  pthread_cond_t cv_work, cv_done;
  int nthreads, started, generation, pending, stop;
  uint8_t* dst;
  const uint8_t* src;
  size_t len;
  pthread_t th[B2_COPY_THREADS_MAX];
  int ids[B2_COPY_THREADS_MAX];
} b2_copy_pool;

static b2_copy_pool g_cp = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, 0, 0, 0, 0, 0, NULL, NULL, 0, {0}, {0}};
static pthread_mutex_t g_cp_user = PTHREAD_MUTEX_INITIALIZER;      /* one parallel copy at a time */

static void* copy_worker(void* arg) {
  const int id = *(int*)arg;
  int seen = 0;
  pthread_mutex_lock(&g_cp.mu);
  for (;;) {
    while (!g_cp.stop && g_cp.generation == seen) pthread_cond_wait(&g_cp.cv_work, &g_cp.mu);
    if (g_cp.stop) break;
    seen = g_cp.generation;
    {
      const size_t per = ((g_cp.len + g_cp.nthreads - 1) / g_cp.nthreads + 63) & ~(size_t)63;
      const size_t lo = per * (size_t)id, hi = lo + per < g_cp.len ? lo + per : g_cp.len;
      uint8_t* d = g_cp.dst;
      const uint8_t* s = g_cp.src;
      pthread_mutex_unlock(&g_cp.mu);
      if (lo < hi) memcpy(d + lo, s + lo, hi - lo);
      pthread_mutex_lock(&g_cp.mu);
    }
    if (--g_cp.pending == 0) pthread_cond_signal(&g_cp.cv_done);
  }
  pthread_mutex_unlock(&g_cp.mu);
  return NULL;
}

static void parallel_memcpy(void* dst, const void* src, size_t len) {
  int i;
  if (len < ((size_t)1 << 20)) { memcpy(dst, src, len); return; }
  pthread_mutex_lock(&g_cp_user);
  pthread_mutex_lock(&g_cp.mu);
  if (!g_cp.started) {
    const char* e = getenv("BLOSC_B200_COPY_THREADS");
    int n = e ? atoi(e) : 8;
    if (n < 1) n = 1;
    if (n > B2_COPY_THREADS_MAX) n = B2_COPY_THREADS_MAX;
    g_cp.nthreads = 0;
    for (i = 0; i < n; i++) {
      g_cp.ids[i] = i;
      if (pthread_create(&g_cp.th[i], NULL, copy_worker, &g_cp.ids[i]) != 0) break;
      g_cp.nthreads++;
    }
    g_cp.started = 1;
  }
  if (g_cp.nthreads == 0) {
    pthread_mutex_unlock(&g_cp.mu);
    memcpy(dst, src, len);
  } else {
    g_cp.dst = (uint8_t*)dst; g_cp.src = (const uint8_t*)src; g_cp.len = len;
    g_cp.pending = g_cp.nthreads;
    g_cp.generation++;
    pthread_cond_broadcast(&g_cp.cv_work);
    while (g_cp.pending) pthread_cond_wait(&g_cp.cv_done, &g_cp.mu);
    pthread_mutex_unlock(&g_cp.mu);
  }
  pthread_mutex_unlock(&g_cp_user);
}

/* ------------------------------------------------------------------------- */
/* workspace pool: device scratch + one stream per concurrent call            */
/* ------------------------------------------------------------------------- */
typedef struct {
  void* p;
  size_t cap;
} b2_buf;

typedef struct {
  int in_use, ready, dev;
  b2_stream_t stream;
  b2_buf in, filt, slots, out, csizes, needs, bstarts;
  b2_buf prev, segs, seg_done, ptail;   /* segment-parallel LZ4 parse (dev_lz4fast.cuh) */
  int* d_result;        /* B2_R_* words (b2_args.h): cbytes, fits, status, work-queue and done counters */
  int* h_result;        /* pinned mirror */
  unsigned queue_base;  /* tickets drawn from the B2_R_QUEUE counter so far (dev_chunk.cuh next_stream) */
  uint8_t* stage[B2_STAGE_DEPTH];   /* pinned bounce slices for pageable host buffers (lazy) */
  b2_event_t stage_ev[B2_STAGE_DEPTH];
  int stage_ok;         /* all DEPTH slices and events exist */
} b2_ws;

#define B2_MAX_WS 16
static b2_ws g_ws[B2_MAX_WS];
static pthread_mutex_t g_ws_mutex = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_ws_cv = PTHREAD_COND_INITIALIZER;
static int g_backend_state = 0;   /* 0 untried, 1 ok, -1 failed */

static int backend_ready(void) {
  int st;
  pthread_mutex_lock(&g_ws_mutex);
  if (g_backend_state == 0) g_backend_state = b2_backend_init() == 0 ? 1 : -1;
  st = g_backend_state;
  pthread_mutex_unlock(&g_ws_mutex);
  if (st < 0) fprintf(stderr, "blosc_b200: no usable CUDA device -- this library has no CPU codec\n");
  return st > 0;
}

static void buf_free(b2_buf* b) { if (b->p) b2_dev_free(b->p); b->p = NULL; b->cap = 0; }

/* Everything a slot owns; the caller holds the slot (in_use) and the slot's device need not be current
 * (device and page-locked allocations are freed by address) */
static void ws_teardown(b2_ws* w) {
  int k;
  buf_free(&w->in); buf_free(&w->filt); buf_free(&w->slots); buf_free(&w->out);
  buf_free(&w->csizes); buf_free(&w->needs); buf_free(&w->bstarts);
  buf_free(&w->prev); buf_free(&w->segs); buf_free(&w->seg_done); buf_free(&w->ptail);
  for (k = 0; k < B2_STAGE_DEPTH; k++) {
    if (w->stage[k]) b2_pinned_free(w->stage[k]);
    if (w->stage_ev[k]) b2_event_destroy(w->stage_ev[k]);
    w->stage[k] = NULL; w->stage_ev[k] = NULL;
  }
  w->stage_ok = 0;
  if (w->d_result) b2_dev_free(w->d_result);
  if (w->h_result) b2_pinned_free(w->h_result);
  if (w->stream) b2_stream_destroy(w->stream);
  w->d_result = NULL; w->h_result = NULL; w->stream = NULL;
  w->ready = 0;
}

/* error paths only: wait for whatever was launched and zero the counter words */
static void ws_reset_counters(b2_ws* w) {
  b2_stream_sync(w->stream);
  b2_memset_dev(w->d_result, 0, 4 * B2_R_WORDS, w->stream);
  if (w->seg_done.p) b2_memset_dev(w->seg_done.p, 0, w->seg_done.cap, w->stream);
  b2_stream_sync(w->stream);
  w->queue_base = 0;
}

/* calls in flight on `dev` right now (the caller's own included) */
static int ws_busy(int dev) {
  int i, n = 0;
  pthread_mutex_lock(&g_ws_mutex);
  for (i = 0; i < B2_MAX_WS; i++) if (g_ws[i].in_use && g_ws[i].ready && g_ws[i].dev == dev) n++;
  pthread_mutex_unlock(&g_ws_mutex);
  return n;
}

static void ws_release(b2_ws* w) {
  pthread_mutex_lock(&g_ws_mutex);
  w->in_use = 0;
  pthread_cond_signal(&g_ws_cv);
  pthread_mutex_unlock(&g_ws_mutex);
}

/* A workspace = one stream + scratch on one device.  The reference's _ctx calls have no limit on the
 * number of concurrent callers (each allocates its own context, blosc.c:1287-1309); here the 17th
 * concurrent call WAITS for a slot instead of failing, and an idle slot that was made on another
 * device is rebuilt for the caller's device when no matching or unused slot is left. */
static b2_ws* ws_acquire(void) {
  b2_ws* w = NULL;
  int i, dev, rebuild = 0;
  if (!backend_ready()) return NULL;
  dev = b2_get_device();
  pthread_mutex_lock(&g_ws_mutex);
  for (;;) {
    for (i = 0; i < B2_MAX_WS && !w; i++)
      if (!g_ws[i].in_use && g_ws[i].ready && g_ws[i].dev == dev) w = &g_ws[i];
    for (i = 0; i < B2_MAX_WS && !w; i++)
      if (!g_ws[i].in_use && !g_ws[i].ready) w = &g_ws[i];
    for (i = 0; i < B2_MAX_WS && !w; i++)
      if (!g_ws[i].in_use) { w = &g_ws[i]; rebuild = 1; }
    if (w) { w->in_use = 1; break; }
    pthread_cond_wait(&g_ws_cv, &g_ws_mutex);
  }
  pthread_mutex_unlock(&g_ws_mutex);
  if (rebuild) ws_teardown(w);
  if (!w->ready) {
    void* p = NULL;
    int ok = 0;
    do {
      if (b2_device_prepare() || b2_stream_create(&w->stream)) break;
      if (b2_dev_alloc(&p, 4 * B2_R_WORDS)) break;
      w->d_result = (int*)p;
      if (b2_pinned_alloc(&p, 4 * B2_R_WORDS)) break;
      w->h_result = (int*)p;
      /* the work counters start at zero and every launch leaves them at zero again (dev_chunk.cuh) */
      if (b2_memset_dev(w->d_result, 0, 4 * B2_R_WORDS, w->stream) || b2_stream_sync(w->stream)) break;
      w->queue_base = 0;
      ok = 1;
    } while (0);
    if (!ok) { ws_teardown(w); ws_release(w); return NULL; }
    w->dev = dev;
    w->ready = 1;
  }
  return w;
}

static int buf_ensure(b2_buf* b, size_t need) {
  if (need <= b->cap) return 0;
  if (b->p) b2_dev_free(b->p);
  b->p = NULL; b->cap = 0;
  need = (need + (need >> 3) + 4095) & ~(size_t)4095;    /* slack so slowly growing sizes do not thrash */
  if (b2_dev_alloc(&b->p, need)) { fprintf(stderr, "blosc_b200: device allocation of %zu bytes failed\n", need); return -1; }
  b->cap = need;
  return 0;
}

/* a buffer of counters that the kernels leave at zero: cleared only when it is (re)allocated */
static int buf_ensure_zeroed(b2_ws* w, b2_buf* b, size_t need) {
  if (need <= b->cap) return 0;
  if (buf_ensure(b, need)) return -1;
  return b2_memset_dev(b->p, 0, b->cap, w->stream);
}

int blosc_free_resources(void) {                              /* blosc.h:411 */
  int i;
  pthread_mutex_lock(&g_ws_mutex);
  for (i = 0; i < B2_MAX_WS; i++) {
    b2_ws* w = &g_ws[i];
    if (w->in_use || !w->ready) continue;
    buf_free(&w->in); buf_free(&w->filt); buf_free(&w->slots); buf_free(&w->out);
    buf_free(&w->csizes); buf_free(&w->needs); buf_free(&w->bstarts);
    buf_free(&w->prev); buf_free(&w->segs); buf_free(&w->seg_done); buf_free(&w->ptail);
  }
  pthread_mutex_unlock(&g_ws_mutex);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* names                                                                      */
/* ------------------------------------------------------------------------- */
int blosc_compcode_to_compname(int compcode, const char** compname) {    /* blosc.c:329-374 */
  static const char* names[6] = {BLOSC_BLOSCLZ_COMPNAME, BLOSC_LZ4_COMPNAME, BLOSC_LZ4HC_COMPNAME,
                                 BLOSC_SNAPPY_COMPNAME, BLOSC_ZLIB_COMPNAME, BLOSC_ZSTD_COMPNAME};
  *compname = (compcode >= 0 && compcode < 6) ? names[compcode] : NULL;
  /* codecs this build can ENCODE; like a reference built without the others */
  if (compcode == BLOSC_BLOSCLZ || compcode == BLOSC_LZ4 || compcode == BLOSC_LZ4HC) return compcode;
  return -1;
}

int blosc_compname_to_compcode(const char* compname) {                    /* blosc.c:377-409 */
  if (strcmp(compname, BLOSC_BLOSCLZ_COMPNAME) == 0) return BLOSC_BLOSCLZ;
  if (strcmp(compname, BLOSC_LZ4_COMPNAME) == 0) return BLOSC_LZ4;
  if (strcmp(compname, BLOSC_LZ4HC_COMPNAME) == 0) return BLOSC_LZ4HC;
  return -1;
}

const char* blosc_list_compressors(void) { return BLOSC_BLOSCLZ_COMPNAME "," BLOSC_LZ4_COMPNAME "," BLOSC_LZ4HC_COMPNAME; }   /* blosc.c:2033-2056 */
const char* blosc_get_version_string(void) { return BLOSC_VERSION_STRING; }

int blosc_get_complib_info(const char* compname, char** complib, char** version) {   /* blosc.c:2063-2124 */
  int code = -1;
  const char *lib = NULL, *ver = NULL;
  if (strcmp(compname, BLOSC_BLOSCLZ_COMPNAME) == 0) { code = BLOSC_BLOSCLZ_LIB; lib = "BloscLZ"; ver = "2.5.1"; }
  else if (strcmp(compname, BLOSC_LZ4_COMPNAME) == 0 || strcmp(compname, BLOSC_LZ4HC_COMPNAME) == 0) {
    code = BLOSC_LZ4_LIB; lib = "LZ4"; ver = "1.10.0";
  }
  if (code < 0) {
    if (complib) *complib = NULL;
    if (version) *version = NULL;
    return -1;
  }
  if (complib) *complib = strdup(lib);
  if (version) *version = strdup(ver);
  return code;
}

static const char* clib_name(int clibcode) {                               /* blosc.c:315-322 */
  static const char* n[5] = {"BloscLZ", "LZ4", "Snappy", "Zlib", "Zstd"};
  return (clibcode >= 0 && clibcode < 5) ? n[clibcode] : NULL;
}

/* ------------------------------------------------------------------------- */
/* header introspection (host pointers, as in the reference)                   */
/* ------------------------------------------------------------------------- */
void blosc_cbuffer_sizes(const void* cbuffer, size_t* nbytes, size_t* cbytes, size_t* blocksize) {   /* blosc.c:2127-2141 */
  const uint8_t* s = (const uint8_t*)cbuffer;
  if (s[0] != BLOSC_VERSION_FORMAT) { *nbytes = *blocksize = *cbytes = 0; return; }
  *nbytes = (size_t)rd_i32(s + 4);
  *blocksize = (size_t)rd_i32(s + 8);
  *cbytes = (size_t)rd_i32(s + 12);
}

int blosc_cbuffer_validate(const void* cbuffer, size_t cbytes, size_t* nbytes) {   /* blosc.c:2143-2150 */
  size_t hc, hb;
  if (cbytes < BLOSC_MIN_HEADER_LENGTH) return -1;
  blosc_cbuffer_sizes(cbuffer, nbytes, &hc, &hb);
  if (hc != cbytes) return -1;
  if (*nbytes > BLOSC_MAX_BUFFERSIZE) return -1;
  return 0;
}

void blosc_cbuffer_metainfo(const void* cbuffer, size_t* typesize, int* flags) {   /* blosc.c:2153-2168 */
  const uint8_t* s = (const uint8_t*)cbuffer;
  if (s[0] != BLOSC_VERSION_FORMAT) { *flags = 0; *typesize = 0; return; }
  *flags = (int)s[2] & 7;
  *typesize = (size_t)s[3];
}

void blosc_cbuffer_versions(const void* cbuffer, int* version, int* versionlz) {   /* blosc.c:2172-2180 */
  const uint8_t* s = (const uint8_t*)cbuffer;
  *version = (int)s[0];
  *versionlz = (int)s[1];
}

const char* blosc_cbuffer_complib(const void* cbuffer) {                            /* blosc.c:2184-2195 */
  const uint8_t* s = (const uint8_t*)cbuffer;
  return clib_name((s[2] & 0xe0) >> 5);
}

/* ------------------------------------------------------------------------- */
/* planning                                                                   */
/* ------------------------------------------------------------------------- */
static int is_hcr(int compcode) { return compcode == BLOSC_LZ4HC || compcode == BLOSC_ZLIB || compcode == BLOSC_ZSTD; }

static int split_block(int compcode, int typesize, int blocksize) {       /* blosc.c:929-959 */
  switch (g_splitmode) {
    case BLOSC_ALWAYS_SPLIT: return 1;
    case BLOSC_NEVER_SPLIT: return 0;
    case BLOSC_AUTO_SPLIT:
      return (compcode == BLOSC_BLOSCLZ || compcode == BLOSC_SNAPPY) && typesize <= MAX_SPLITS &&
             blocksize / typesize >= MIN_BUFFERSIZE;
    case BLOSC_FORWARD_COMPAT_SPLIT:
      return compcode != BLOSC_ZSTD && typesize <= MAX_SPLITS && blocksize / typesize >= MIN_BUFFERSIZE;
    default:
      fprintf(stderr, "Split mode %d not supported", g_splitmode);
      return -1;
  }
}

static int32_t compute_blocksize(int compcode, int clevel, int32_t typesize, int32_t nbytes, int32_t forced) {   /* blosc.c:962-1060 */
  int32_t bs = nbytes;
  if (nbytes < typesize) return 1;
  if (forced) {
    bs = forced;
    if (bs < MIN_BUFFERSIZE) bs = MIN_BUFFERSIZE;
    if (bs > (int32_t)BLOSC_MAX_BLOCKSIZE) bs = (int32_t)BLOSC_MAX_BLOCKSIZE;
  } else if (nbytes >= L1_SIZE) {
    bs = L1_SIZE;
    if (is_hcr(compcode)) bs *= 2;
    switch (clevel) {
      case 0: bs /= 4; break;
      case 1: bs /= 2; break;
      case 2: break;
      case 3: bs *= 2; break;
      case 4: case 5: bs *= 4; break;
      case 6: case 7: case 8: bs *= 8; break;
      default: bs *= 8; if (is_hcr(compcode)) bs *= 2; break;
    }
  }
  if (clevel > 0 && split_block(compcode, typesize, bs)) {
    if (bs > (1 << 18)) bs = 1 << 18;
    bs *= typesize;
    if (bs < (1 << 16)) bs = 1 << 16;
    if (bs > 1024 * 1024) bs = 1024 * 1024;
  }
  if (bs > nbytes) bs = nbytes;
  if (bs > typesize) bs = bs / typesize * typesize;
  return bs;
}

static int check_threads(int nthreads, int32_t nbytes, int32_t blocksize) {
  /* do_job takes the pool path (and so validates nthreads, blosc.c:1977-1986) only when
   * nthreads != 1 and the buffer has more than one block (blosc.c:910) */
  if (nthreads == 1 || nbytes / blocksize <= 1) return 0;
  if (nthreads > BLOSC_MAX_THREADS) {
    fprintf(stderr, "Error.  nthreads cannot be larger than BLOSC_MAX_THREADS (%d)", BLOSC_MAX_THREADS);
    return -1;
  }
  if (nthreads <= 0) { fprintf(stderr, "Error.  nthreads must be a positive integer"); return -1; }
  return 0;
}

/* copy between any combination of host / device memory */
static int copy_any(void* dst, int dst_dev, const void* src, int src_dev, size_t n, b2_stream_t s) {
  int rc = 0;
  if (n == 0) return 0;
  if (!dst_dev && !src_dev) { memcpy(dst, src, n); return 0; }
  if (dst_dev && src_dev) rc = b2_copy_d2d(dst, src, n, s);
  else if (dst_dev) rc = b2_copy_h2d(dst, src, n, s);
  else rc = b2_copy_d2h(dst, src, n, s);
  if (rc == 0) rc = b2_stream_sync(s);
  return rc;
}

static int stage_ready(b2_ws* w) {
  int k;
  if (w->stage_ok) return 0;
  for (k = 0; k < B2_STAGE_DEPTH; k++) {
    void* p = NULL;
    if (!w->stage[k]) { if (b2_pinned_alloc(&p, B2_STAGE_SLICE)) break; w->stage[k] = (uint8_t*)p; }
    if (!w->stage_ev[k] && b2_event_create(&w->stage_ev[k])) { w->stage_ev[k] = NULL; break; }
  }
  if (k < B2_STAGE_DEPTH) {                 /* partial: give everything back, callers fall back to a direct copy */
    for (k = 0; k < B2_STAGE_DEPTH; k++) {
      if (w->stage[k]) b2_pinned_free(w->stage[k]);
      if (w->stage_ev[k]) b2_event_destroy(w->stage_ev[k]);
      w->stage[k] = NULL; w->stage_ev[k] = NULL;
    }
    return -1;
  }
  w->stage_ok = 1;
  return 0;
}

/* host -> device; pageable sources go through the bounce slices */
static int h2d_any(b2_ws* w, void* dst, const void* src, size_t n) {
  size_t off;
  int i = 0;
  if (n == 0) return 0;
  if (n < B2_STAGE_SLICE / 4 || b2_ptr_is_pinned(src) || stage_ready(w)) return b2_copy_h2d(dst, src, n, w->stream);
  for (off = 0; off < n; off += B2_STAGE_SLICE, i++) {
    const int k = i % B2_STAGE_DEPTH;
    const size_t len = n - off < B2_STAGE_SLICE ? n - off : B2_STAGE_SLICE;
    if (i >= B2_STAGE_DEPTH && b2_event_sync(w->stage_ev[k])) return -1;     /* slice k's previous DMA has drained */
    parallel_memcpy(w->stage[k], (const uint8_t*)src + off, len);
    if (b2_copy_h2d((uint8_t*)dst + off, w->stage[k], len, w->stream)) return -1;
    if (b2_event_record(w->stage_ev[k], w->stream)) return -1;
  }
  return 0;
}

/* device -> host, completes before returning; pageable destinations go through the bounce slices */
static int d2h_any(b2_ws* w, void* dst, const void* src, size_t n) {
  size_t off;
  int i = 0, nsl;
  if (n == 0) return 0;
  if (n < B2_STAGE_SLICE / 4 || b2_ptr_is_pinned(dst) || stage_ready(w)) {
    if (b2_copy_d2h(dst, src, n, w->stream)) return -1;
    return b2_stream_sync(w->stream);
  }
  nsl = (int)((n + B2_STAGE_SLICE - 1) / B2_STAGE_SLICE);
  for (i = 0; i < nsl + B2_STAGE_DEPTH - 1; i++) {
    if (i < nsl) {                                   /* issue the DMA of slice i */
      const int k = i % B2_STAGE_DEPTH;
      off = (size_t)i * B2_STAGE_SLICE;
      if (b2_copy_d2h(w->stage[k], (const uint8_t*)src + off, n - off < B2_STAGE_SLICE ? n - off : B2_STAGE_SLICE, w->stream)) return -1;
      if (b2_event_record(w->stage_ev[k], w->stream)) return -1;
    }
    if (i >= B2_STAGE_DEPTH - 1) {                   /* drain slice j = i - (DEPTH-1) to the caller's buffer */
      const int j = i - (B2_STAGE_DEPTH - 1), k = j % B2_STAGE_DEPTH;
      off = (size_t)j * B2_STAGE_SLICE;
      if (b2_event_sync(w->stage_ev[k])) return -1;
      parallel_memcpy((uint8_t*)dst + off, w->stage[k], n - off < B2_STAGE_SLICE ? n - off : B2_STAGE_SLICE);
    }
  }
  return 0;
}

static void make_header(uint8_t* h, int versionlz, int flags, int typesize, int32_t nbytes, int32_t blocksize,
                        int32_t cbytes) {                                  /* blosc.c:1154-1215,1275 */
  h[0] = BLOSC_VERSION_FORMAT; h[1] = (uint8_t)versionlz; h[2] = (uint8_t)flags; h[3] = (uint8_t)typesize;
  wr_i32(h + 4, nbytes); wr_i32(h + 8, blocksize); wr_i32(h + 12, cbytes);
}

/* Where a finished chunk goes.  The plain API writes at `dest`; a frame (below) learns the
 * chunk's offset only once every earlier chunk has announced its size, so the compressor asks
 * for the final location at the moment `cbytes` is known. */
typedef struct b2_frame_job b2_frame_job;
typedef struct {
  b2_frame_job* job;
  int index, placed;
  int many;             /* other chunks are in flight with this one: favour streams per SM over latency */
} b2_place;
static void* frame_place(b2_place* pl, int32_t cbytes);

/* Opt-in (BLOSC_B200_LZ4_PACK=1): measured on the 8 GiB frames workload it helps typesize 2 (+20 %)
 * and 8 (+6 %) but costs 12 % at typesize 4, where the extra instructions per probe outweigh the
 * doubled number of streams per SM -- so it is not the default. */
static int lz4_pack_wanted(const b2_place* pl) {
  const char* e = getenv("BLOSC_B200_LZ4_PACK");
  (void)pl;
  return e && *e && atoi(e) != 0;
}

/* BLOSC_B200_PARSE selects the LZ4 encoder: "exact" (default) replays LZ4_compress_fast bit for bit, chunks are
 * byte-identical to the reference's; "fast" (alias "segmented") is the segment-parallel parser of dev_lz4fast.cuh:
 * chunks are valid Blosc-1 / LZ4 that any reference build decodes, but not the reference's bytes. */
static int lz4_fast_wanted(void) {
  const char* e = getenv("BLOSC_B200_PARSE");
  return e && (strcmp(e, "fast") == 0 || strcmp(e, "segmented") == 0);
}

/* header + raw payload (blosc.c:825-830) */
static int emit_memcpyed(const uint8_t* hdr, const void* src, int src_dev, void* dest, int dest_dev, int32_t nbytes) {
  b2_ws* w = NULL;
  int rc = 0;
  if (src_dev || dest_dev) { w = ws_acquire(); if (!w) return -1; }
  rc = copy_any(dest, dest_dev, hdr, 0, 16, w ? w->stream : NULL);
  if (!rc) rc = copy_any((uint8_t*)dest + 16, dest_dev, src, src_dev, (size_t)nbytes, w ? w->stream : NULL);
  if (w) ws_release(w);
  return rc ? -1 : nbytes + 16;
}

/* ------------------------------------------------------------------------- */
/* compression                                                                */
/* ------------------------------------------------------------------------- */
static int compress_impl(int clevel, int doshuffle, size_t typesize, size_t nbytes, const void* src, void* dest,
                         size_t destsize, const char* compressor, size_t blocksize, int numinternalthreads,
                         b2_place* pl, int pl_dest_dev) {
  const int compcode = blosc_compname_to_compcode(compressor);
  int32_t ts, nb, bs, nblocks, leftover, dsz;
  int flags = 0, compformat, dont_split, src_dev, dest_dev, dofilter, fmode = 0, nsplits, result = -1, launched = 0;
  uint8_t hdr[16];
  b2_ws* w;
  const uint8_t* d_src;
  const uint8_t* d_codec_in;
  uint8_t* d_dest;

  /* initialize_context_compression, blosc.c:1062-1145 */
  if (nbytes > BLOSC_MAX_BUFFERSIZE) return 0;
  if (destsize < BLOSC_MAX_OVERHEAD) return 0;
  if (destsize - BLOSC_MAX_OVERHEAD > nbytes) destsize = nbytes + BLOSC_MAX_OVERHEAD;
  if (clevel < 0 || clevel > 9) return -10;
  if (doshuffle != 0 && doshuffle != 1 && doshuffle != 2) return -10;
  if (typesize == 0) return -10;
  if (typesize > BLOSC_MAX_TYPESIZE) typesize = 1;
  ts = (int32_t)typesize; nb = (int32_t)nbytes; dsz = (int32_t)destsize;
  bs = compute_blocksize(compcode, clevel, ts, nb, (int32_t)blocksize);
  nblocks = nb / bs; leftover = nb % bs;
  if (leftover > 0) nblocks++;

  /* write_compression_header, blosc.c:1148-1247 */
  if (compcode == BLOSC_BLOSCLZ) compformat = BLOSC_BLOSCLZ_FORMAT;
  else if (compcode == BLOSC_LZ4) compformat = BLOSC_LZ4_FORMAT;
  else if (compcode == BLOSC_LZ4HC) compformat = BLOSC_LZ4HC_FORMAT;     /* blosc.c:1170-1172: the LZ4 format */
  else {
    fprintf(stderr, "Blosc has not been compiled with '%s' ", compressor ? compressor : "(null)");
    fprintf(stderr, "compression support.  Please use one having it.");
    return -5;
  }
  if (clevel == 0) flags |= BLOSC_MEMCPYED;
  if (nb < MIN_BUFFERSIZE) flags |= BLOSC_MEMCPYED;
  if (doshuffle == BLOSC_SHUFFLE) flags |= BLOSC_DOSHUFFLE;
  if (doshuffle == BLOSC_BITSHUFFLE) flags |= BLOSC_DOBITSHUFFLE;
  dont_split = !split_block(compcode, ts, bs);
  flags |= dont_split << 4;
  flags |= compformat << 5;

  src_dev = (nb > 0) ? b2_ptr_is_device(src) : 0;
  dest_dev = pl ? pl_dest_dev : b2_ptr_is_device(dest);

  /* blosc_compress_context, blosc.c:1250-1279 */
  if ((flags & BLOSC_MEMCPYED) && nb + BLOSC_MAX_OVERHEAD > dsz) return 0;
  if (check_threads(numinternalthreads, nb, bs) < 0) return -1;
  if (flags & BLOSC_MEMCPYED) {
    make_header(hdr, 1, flags, ts, nb, bs, nb + 16);
    if (pl && !(dest = frame_place(pl, nb + 16))) return 0;
    return emit_memcpyed(hdr, src, src_dev, dest, dest_dev, nb);
  }

  w = ws_acquire();
  if (!w) return -1;
  do {
    FilterArgs fa;
    EncodeArgs ea;
    ScanArgs sa;
    CompactArgs ca;
    const int32_t nfull = nb / bs;
    /* stage the input on the device if it lives in host memory */
    if (src_dev) d_src = (const uint8_t*)src;
    else {
      if (buf_ensure(&w->in, (size_t)nb + 64)) break;
      if (h2d_any(w, w->in.p, src, (size_t)nb)) break;
      d_src = (const uint8_t*)w->in.p;
    }
    /* filter (blosc.c:607-622): byte shuffle needs typesize > 1; bitshuffle applies per block when bsize >= typesize */
    dofilter = ((flags & BLOSC_DOSHUFFLE) && ts > 1) || (flags & BLOSC_DOBITSHUFFLE);
    if ((flags & BLOSC_DOSHUFFLE) && ts > 1) fmode = FILT_SHUFFLE; else fmode = FILT_BITSHUFFLE;
    d_codec_in = d_src;
    if (dofilter) {
      if (buf_ensure(&w->filt, (size_t)nb + 64)) break;
      fa.src = d_src; fa.dst = (uint8_t*)w->filt.p; fa.nbytes = nb; fa.blocksize = bs; fa.typesize = ts; fa.mode = fmode;
      if (b2_launch_filter(&fa, w->stream)) break;
      d_codec_in = (const uint8_t*)w->filt.p;
    }
    /* one LZ stream per split (blosc.c:628-634) */
    nsplits = dont_split ? 1 : ts;
    memset(&ea, 0, sizeof ea);
    ea.map.nbytes = nb; ea.map.blocksize = bs; ea.map.nsplits = nsplits; ea.map.first_block = 0;
    ea.map.nfull = nfull; ea.map.leftover = leftover; ea.map.nstreams = nfull * nsplits + (leftover ? 1 : 0);
    if (buf_ensure(&w->slots, (size_t)nb + 64)) break;
    if (buf_ensure(&w->csizes, (size_t)ea.map.nstreams * 4 + 64)) break;
    if (buf_ensure(&w->needs, (size_t)ea.map.nstreams * 4 + 64)) break;
    if (buf_ensure(&w->bstarts, (size_t)nblocks * 4 + 64)) break;
    ea.in = d_codec_in; ea.slots = (uint8_t*)w->slots.p; ea.csizes = (int*)w->csizes.p; ea.needs = (int*)w->needs.p;
    ea.codec = (compcode == BLOSC_LZ4 || compcode == BLOSC_LZ4HC) ? B2_CODEC_LZ4 : B2_CODEC_BLOSCLZ;
    ea.clevel = clevel; ea.accel = 10 - clevel;                          /* blosc.c:577-587 */
    ea.split_flag = !dont_split;
    ea.many = (pl && pl->many) || ws_busy(w->dev) > 1;      /* a frame, or other _ctx calls running on this device */
    ea.table_bytes = ea.codec == B2_CODEC_LZ4 ? 16384 : (4 << (clevel == 1 ? 12 : (clevel == 2 ? 13 : 14)));
    /* BloscLZ at clevel >= 3: 17-bit packed table (34 KiB instead of 64 KiB) when every stream is <= 128 KiB */
    if (ea.codec == B2_CODEC_BLOSCLZ && clevel >= 3 && bs / nsplits <= 131072 && leftover <= 131072) ea.table_bytes = 32768 + 2048;
    /* LZ4, on request: 17-bit packed table, 8.5 KiB instead of 16 KiB per stream (twice the streams per SM),
     * when every stream is long enough for the 4096-entry table (lz4.c:710) and at most 128 KiB */
    if (ea.codec == B2_CODEC_LZ4 && lz4_pack_wanted(pl) && leftover == 0 && bs / nsplits >= 65547 && bs / nsplits <= 131072)
      ea.table_bytes = 8192 + 512;
    ea.queue = w->d_result + B2_R_QUEUE; ea.queue_base_host = &w->queue_base; ea.done = w->d_result + B2_R_DONE;
    sa.csizes = ea.csizes; sa.needs = ea.needs; sa.blocksize = bs; sa.leftover = leftover;
    /* do_job runs serial_blosc when nthreads == 1 or there is at most one block (blosc.c:910) */
    sa.serial = (numinternalthreads == 1 || nb / bs <= 1);
    sa.bstarts = (int*)w->bstarts.p; sa.result = w->d_result;
    sa.nsplits = nsplits; sa.nfull = nfull; sa.has_leftover = leftover > 0; sa.destsize = dsz;
    /* the warp that finishes the last stream also does the block scan (no separate 1-CTA launch) */
    ea.fold_scan = nblocks <= B2_FOLD_SCAN_MAX_BLOCKS;
    ea.scan = sa;
    launched = 1;
    memset(&ca, 0, sizeof ca);
    if (ea.codec == B2_CODEC_LZ4 && (compcode == BLOSC_LZ4HC || lz4_fast_wanted())) {
      FastArgs fx;
      const int neblock = bs / nsplits;
      memset(&fx, 0, sizeof fx);
      fx.map = ea.map; fx.in = ea.in; fx.slots = ea.slots; fx.csizes = ea.csizes; fx.needs = ea.needs;
      fx.segs_full = (neblock + B2_FAST_SEG - 1) / B2_FAST_SEG; fx.segs_left = (leftover + B2_FAST_SEG - 1) / B2_FAST_SEG;
      {
        /* one parse CTA per window: the whole stream when it fits in B2_FAST_WIN_MAX, rounded up to whole warps of segments */
        const int longest = neblock > leftover ? neblock : leftover;
        int win = (longest + 32 * B2_FAST_SEG - 1) / (32 * B2_FAST_SEG) * (32 * B2_FAST_SEG);
        if (win > B2_FAST_WIN_MAX) win = B2_FAST_WIN_MAX;
        fx.win_bytes = win;
        fx.threads = win / B2_FAST_SEG;
        fx.groups_full = (neblock + win - 1) / win; fx.groups_left = (leftover + win - 1) / win;
      }
      fx.depth = 3 * clevel + 1; fx.accel = ea.accel;
      /* "lz4hc" (blosc.c:422-433 hands clevel to LZ4_compress_HC): the same hash-chain parser, with LZ4HC's search
       * effort -- 2^(level-1) candidates, capped -- and no skipping over literals */
      fx.hash_mask = 0xffff; fx.lazy = 0;
      if (compcode == BLOSC_LZ4HC) {
        fx.depth = clevel <= 2 ? 4 : (clevel >= 8 ? 128 : (1 << (clevel - 1))); fx.accel = 1;
        fx.hash_mask = 0xffff; fx.lazy = 64;
      }
      if (buf_ensure(&w->prev, 2 * (size_t)nb + 64)) break;
      if (buf_ensure(&w->segs, ((size_t)nfull * nsplits * fx.segs_full + fx.segs_left + 8) * sizeof(FastSeg))) break;
      if (buf_ensure_zeroed(w, &w->seg_done, (size_t)ea.map.nstreams * 4 + 64)) break;
      if (buf_ensure(&w->ptail, (size_t)ea.map.nstreams * 4 + 64)) break;
      fx.prev = (uint16_t*)w->prev.p; fx.segs = (FastSeg*)w->segs.p; fx.seg_done = (int*)w->seg_done.p; fx.ptail = (int*)w->ptail.p;
      fx.queue = ea.queue; fx.queue_base_host = ea.queue_base_host; fx.done = ea.done;
      fx.fold_scan = ea.fold_scan; fx.scan = sa;
      if (b2_launch_fast(&fx, w->stream)) break;
      ca.segs = fx.segs; ca.ptail = fx.ptail; ca.segs_full = fx.segs_full; ca.segs_left = fx.segs_left;
    } else if (b2_launch_encode(&ea, w->stream)) break;
    if (!ea.fold_scan && b2_launch_scan(&sa, w->stream)) break;
    if (dest_dev && !pl) d_dest = (uint8_t*)dest;
    else { if (buf_ensure(&w->out, (size_t)dsz + 64)) break; d_dest = (uint8_t*)w->out.p; }
    ca.map = ea.map; ca.in = d_codec_in; ca.slots = ea.slots; ca.csizes = ea.csizes; ca.bstarts = sa.bstarts;
    ca.result = w->d_result; ca.dest = d_dest;
    ca.hdr0 = (uint32_t)BLOSC_VERSION_FORMAT | (1u << 8) | ((uint32_t)flags << 16) | ((uint32_t)ts << 24);
    ca.nbytes32 = nb; ca.nblocks = nblocks;
    if (b2_launch_compact(&ca, w->stream)) break;
    if (b2_copy_d2h(w->h_result, w->d_result, 8, w->stream)) break;
    if (b2_stream_sync(w->stream)) break;
    if (w->h_result[1]) {                                                 /* fits */
      const int32_t cbytes = w->h_result[0];
      if (pl && !(dest = frame_place(pl, cbytes))) { result = 0; break; }
      if (!dest_dev) {
        if (d2h_any(w, dest, d_dest, (size_t)cbytes)) break;
      } else if (pl) {
        if (copy_any(dest, 1, d_dest, 1, (size_t)cbytes, w->stream)) break;
      }
      result = cbytes;
    } else if (nb + BLOSC_MAX_OVERHEAD <= dsz) {                          /* blosc.c:1264-1272 */
      result = -2;   /* marker: redo as MEMCPYED after releasing the workspace */
    } else {
      make_header(hdr, 1, flags, ts, nb, bs, 0);                          /* blosc.c:1275 with ntbytes == 0 */
      if (!pl && copy_any(dest, dest_dev, hdr, 0, 16, w->stream)) break;
      result = 0;
    }
  } while (0);
  if (result == -1 && launched) ws_reset_counters(w);                      /* a failed call must not leave the counters dirty */
  ws_release(w);
  if (result == -2) {
    flags |= BLOSC_MEMCPYED;
    make_header(hdr, 1, flags, ts, nb, bs, nb + 16);
    if (pl && !(dest = frame_place(pl, nb + 16))) return 0;
    return emit_memcpyed(hdr, src, src_dev, dest, dest_dev, nb);
  }
  return result;
}

int blosc_compress_ctx(int clevel, int doshuffle, size_t typesize, size_t nbytes, const void* src, void* dest,
                       size_t destsize, const char* compressor, size_t blocksize, int numinternalthreads) {
  return compress_impl(clevel, doshuffle, typesize, nbytes, src, dest, destsize, compressor, blocksize,
                       numinternalthreads, NULL, 0);
}

/* ------------------------------------------------------------------------- */
/* decompression                                                              */
/* ------------------------------------------------------------------------- */
typedef struct {
  int version, versionlz, flags, typesize;
  int32_t nbytes, blocksize, cbytes, nblocks, leftover;
} b2_hdr;

static void parse_header(const uint8_t* h, b2_hdr* o) {                    /* blosc.c:1453-1461 */
  o->version = h[0]; o->versionlz = h[1]; o->flags = h[2]; o->typesize = h[3];
  o->nbytes = rd_i32(h + 4); o->blocksize = rd_i32(h + 8); o->cbytes = rd_i32(h + 12);
  o->nblocks = 0; o->leftover = 0;
}

static int codec_from_header(const b2_hdr* h, int* codec) {                /* blosc.c:525-574 */
  const int fmt = (h->flags & 0xe0) >> 5;
  if (fmt == BLOSC_BLOSCLZ_FORMAT) { if (h->versionlz != BLOSC_BLOSCLZ_VERSION_FORMAT) return -9; *codec = B2_CODEC_BLOSCLZ; return 0; }
  if (fmt == BLOSC_LZ4_FORMAT) { if (h->versionlz != BLOSC_LZ4_VERSION_FORMAT) return -9; *codec = B2_CODEC_LZ4; return 0; }
  if (fmt == BLOSC_ZLIB_FORMAT) { if (h->versionlz != BLOSC_ZLIB_VERSION_FORMAT) return -9; *codec = B2_CODEC_ZLIB; return 0; }   /* blosc.c:556-561 */
  if (fmt == BLOSC_ZSTD_FORMAT) { if (h->versionlz != BLOSC_ZSTD_VERSION_FORMAT) return -9; *codec = B2_CODEC_ZSTD; return 0; }   /* blosc.c:565-571 */
  return -5;
}

/* Decode blocks [first, first+count) of a chunk that is already on the device into `d_out`
 * (device), which represents buffer offsets [first*blocksize, ...).  Shared by decompress and getitem. */
static int decode_blocks(b2_ws* w, const b2_hdr* h, int codec, const uint8_t* d_chunk, int first, int count,
                         uint8_t* d_out) {
  DecodeArgs da;
  FilterArgs fa;
  const int ts = h->typesize, bs = h->blocksize;
  const int dont_split = (h->flags & 0x10) >> 4;
  const int doshuffle = (h->flags & BLOSC_DOSHUFFLE) && ts > 1;
  const int dobitshuffle = !doshuffle && (h->flags & BLOSC_DOBITSHUFFLE);
  const int has_left = h->leftover > 0 && first + count == h->nblocks;
  const int nfull = count - has_left;
  const long long span = (long long)nfull * bs + (has_left ? h->leftover : 0);
  uint8_t* d_codec_out = d_out;
  memset(&da, 0, sizeof da);
  /* blosc.c:749-757: split only if typesize <= 16 and >= 128 elements per block */
  da.map.nsplits = (!dont_split && ts <= MAX_SPLITS && bs / ts >= MIN_BUFFERSIZE) ? ts : 1;
  da.map.nbytes = h->nbytes; da.map.blocksize = bs; da.map.first_block = first; da.map.nfull = nfull;
  da.map.leftover = has_left ? h->leftover : 0;
  da.map.nstreams = nfull * da.map.nsplits + (has_left ? 1 : 0);
  if (doshuffle || dobitshuffle) {
    if (buf_ensure(&w->filt, (size_t)span + 64)) return -1;
    d_codec_out = (uint8_t*)w->filt.p;
  }
  da.chunk = d_chunk; da.cbytes = h->cbytes; da.out = d_codec_out; da.out_shift = (long long)first * bs;
  da.codec = codec; da.status = w->d_result + B2_R_STATUS; da.queue = w->d_result + B2_R_QUEUE;
  da.queue_base_host = &w->queue_base;
  da.done = w->d_result + B2_R_DONE; da.status_out = w->d_result + B2_R_STATUS_OUT;
  da.many = ws_busy(w->dev) > 1;
  if (b2_launch_decode(&da, w->stream)) { ws_reset_counters(w); return -1; }
  if (doshuffle || dobitshuffle) {
    fa.src = d_codec_out; fa.dst = d_out; fa.nbytes = span; fa.blocksize = bs; fa.typesize = ts;
    fa.mode = doshuffle ? FILT_UNSHUFFLE : FILT_BITUNSHUFFLE;
    if (b2_launch_filter(&fa, w->stream)) { ws_reset_counters(w); return -1; }
  }
  if (b2_copy_d2h(w->h_result + B2_R_STATUS_OUT, w->d_result + B2_R_STATUS_OUT, 4, w->stream) || b2_stream_sync(w->stream)) {
    ws_reset_counters(w);
    return -1;
  }
  return w->h_result[B2_R_STATUS_OUT] < 0 ? w->h_result[B2_R_STATUS_OUT] : 0;
}

/* max_cbytes >= 0 (frames): the chunk lives in a slot of that many bytes and must decode to exactly
 * expect_nbytes -- a chunk header that claims more is refused before anything is copied */
static int decompress_impl(const void* src, void* dest, size_t destsize, int numinternalthreads, long long max_cbytes,
                           long long expect_nbytes) {
  uint8_t hb[16];
  b2_hdr h;
  int src_dev, dest_dev, codec = 0, rc, result = -1;
  b2_ws* w;

  src_dev = b2_ptr_is_device(src);
  if (src_dev) {
    w = ws_acquire();
    if (!w) return -1;
    rc = copy_any(hb, 0, src, 1, 16, w->stream);
    ws_release(w);
    if (rc) return -1;
  } else memcpy(hb, src, 16);
  parse_header(hb, &h);
  if (max_cbytes >= 0 && (h.cbytes < BLOSC_MAX_OVERHEAD || h.cbytes > max_cbytes || h.nbytes != expect_nbytes)) return -1;

  /* blosc_run_decompression_with_context, blosc.c:1463-1508 */
  if (h.nbytes == 0) return 0;
  if (h.blocksize <= 0 || (size_t)h.blocksize > destsize || (size_t)h.blocksize > BLOSC_MAX_BLOCKSIZE || h.typesize <= 0)
    return -1;
  if (h.version != BLOSC_VERSION_FORMAT) return -1;
  if (h.flags & 0x08) return -1;
  h.nblocks = h.nbytes / h.blocksize; h.leftover = h.nbytes % h.blocksize;
  if (h.leftover > 0) h.nblocks++;
  if (h.nbytes > (int32_t)destsize) return -1;
  dest_dev = b2_ptr_is_device(dest);
  if (h.flags & BLOSC_MEMCPYED) {
    if (h.nbytes + BLOSC_MAX_OVERHEAD != h.cbytes) return -1;
  } else {
    rc = codec_from_header(&h, &codec);
    if (rc) return rc;
    if (h.nblocks > (h.cbytes - 16) / 4) return -1;
  }
  /* A negative header nbytes passes every check above in the reference too; its block loop then runs
   * zero times (nblocks <= 0, blosc.c:815, :910) and the call returns 0 without touching memory. */
  if (h.nblocks <= 0) return 0;
  if (check_threads(numinternalthreads, h.nbytes, h.blocksize) < 0) return -1;

  if (h.flags & BLOSC_MEMCPYED) {                                          /* blosc.c:843-848 */
    w = NULL;
    if (src_dev || dest_dev) { w = ws_acquire(); if (!w) return -1; }
    rc = copy_any(dest, dest_dev, (const uint8_t*)src + 16, src_dev, (size_t)h.nbytes, w ? w->stream : NULL);
    if (w) ws_release(w);
    return rc ? -1 : h.nbytes;
  }

  w = ws_acquire();
  if (!w) return -1;
  do {
    const uint8_t* d_chunk;
    uint8_t* d_out;
    if (src_dev) d_chunk = (const uint8_t*)src;
    else {
      if (buf_ensure(&w->in, (size_t)h.cbytes + 64)) break;
      if (h2d_any(w, w->in.p, src, (size_t)h.cbytes)) break;
      d_chunk = (const uint8_t*)w->in.p;
    }
    if (dest_dev) d_out = (uint8_t*)dest;
    else { if (buf_ensure(&w->out, (size_t)h.nbytes + 64)) break; d_out = (uint8_t*)w->out.p; }
    rc = decode_blocks(w, &h, codec, d_chunk, 0, h.nblocks, d_out);
    if (rc < 0) { result = -1; break; }                                    /* blosc.c:1511-1514 */
    if (!dest_dev) {
      if (d2h_any(w, dest, d_out, (size_t)h.nbytes)) break;
    }
    result = h.nbytes;
  } while (0);
  ws_release(w);
  return result;
}

int blosc_decompress_ctx(const void* src, void* dest, size_t destsize, int numinternalthreads) {
  return decompress_impl(src, dest, destsize, numinternalthreads, -1, -1);
}

static int getitem_impl(const void* src, int start, int nitems, void* dest, long long max_cbytes) {    /* blosc.c:1574-1703 */
  uint8_t hb[16];
  b2_hdr h;
  int src_dev, dest_dev, codec = 0, rc, result = -1;
  const int stop = start + nitems;
  b2_ws* w;
  long long b_lo, b_hi, first, last;

  src_dev = b2_ptr_is_device(src);
  if (src_dev) {
    w = ws_acquire();
    if (!w) return -1;
    rc = copy_any(hb, 0, src, 1, 16, w->stream);
    ws_release(w);
    if (rc) return -1;
  } else memcpy(hb, src, 16);
  parse_header(hb, &h);
  if (max_cbytes >= 0 && (h.cbytes < BLOSC_MAX_OVERHEAD || h.cbytes > max_cbytes)) return -1;
  if (h.version != BLOSC_VERSION_FORMAT) return -9;
  if (h.blocksize <= 0 || h.blocksize > h.nbytes || (size_t)h.blocksize > BLOSC_MAX_BLOCKSIZE || h.typesize <= 0)
    return -1;
  h.nblocks = h.nbytes / h.blocksize; h.leftover = h.nbytes % h.blocksize;
  if (h.leftover > 0) h.nblocks++;
  if (h.flags & BLOSC_MEMCPYED) {
    if (h.nbytes + BLOSC_MAX_OVERHEAD != h.cbytes) return -1;
  } else {
    rc = codec_from_header(&h, &codec);
    if (rc) return rc;
    if (h.nblocks >= (h.cbytes - 16) / 4) return -1;                       /* :1630 */
  }
  if (start < 0 || (long long)start * h.typesize > h.nbytes) { fprintf(stderr, "`start` out of bounds"); return -1; }
  if (stop < 0 || (long long)stop * h.typesize > h.nbytes) { fprintf(stderr, "`start`+`nitems` out of bounds"); return -1; }
  b_lo = (long long)start * h.typesize; b_hi = (long long)stop * h.typesize;
  if (b_hi <= b_lo) return 0;                                              /* no block overlaps: loop copies nothing */
  dest_dev = b2_ptr_is_device(dest);

  if (h.flags & BLOSC_MEMCPYED) {                                          /* :1678-1683 */
    w = NULL;
    if (src_dev || dest_dev) { w = ws_acquire(); if (!w) return -1; }
    rc = copy_any(dest, dest_dev, (const uint8_t*)src + 16 + b_lo, src_dev, (size_t)(b_hi - b_lo), w ? w->stream : NULL);
    if (w) ws_release(w);
    return rc ? -1 : (int)(b_hi - b_lo);
  }

  first = b_lo / h.blocksize; last = (b_hi - 1) / h.blocksize;             /* only overlapping blocks are decoded, :1666 */
  w = ws_acquire();
  if (!w) return -1;
  do {
    const uint8_t* d_chunk;
    const int count = (int)(last - first + 1);
    if (src_dev) d_chunk = (const uint8_t*)src;
    else {
      /* only what the decoder will read crosses PCIe: the header with bstarts[], and the bytes of the blocks that
       * overlap the request -- from the lowest of their bstarts to the next bstart above the highest (blosc.c:1655-1668
       * decodes just those blocks; the kernel bounds-checks every offset against cbytes as blosc_d does) */
      const uint8_t* hs = (const uint8_t*)src;
      const size_t index_end = 16 + 4 * (size_t)h.nblocks;
      size_t lo = (size_t)h.cbytes, hi = 0, cut = (size_t)h.cbytes;
      long long b;
      int bad = 0;
      if (buf_ensure(&w->in, (size_t)h.cbytes + 64)) break;
      for (b = first; b <= last; b++) {
        const int32_t bs_b = rd_i32(hs + 16 + 4 * b);
        if (bs_b < (int32_t)index_end || bs_b > h.cbytes) { bad = 1; break; }
        if ((size_t)bs_b < lo) lo = (size_t)bs_b;
        if ((size_t)bs_b > hi) hi = (size_t)bs_b;
      }
      if (bad) { result = -1; break; }                                     /* blosc_d would refuse this offset (blosc.c:761) */
      for (b = 0; b < h.nblocks; b++) {                                    /* where the last needed block ends */
        const int32_t bs_b = rd_i32(hs + 16 + 4 * b);
        if ((size_t)bs_b > hi && (size_t)bs_b < cut) cut = (size_t)bs_b;
      }
      if (h2d_any(w, w->in.p, hs, index_end)) break;
      if (h2d_any(w, (uint8_t*)w->in.p + lo, hs + lo, cut - lo)) break;
      d_chunk = (const uint8_t*)w->in.p;
    }
    if (buf_ensure(&w->out, (size_t)count * (size_t)h.blocksize + 64)) break;
    rc = decode_blocks(w, &h, codec, d_chunk, (int)first, count, (uint8_t*)w->out.p);
    if (rc < 0) { result = rc; break; }                                    /* :1689-1692 returns blosc_d's code */
    if (copy_any(dest, dest_dev, (const uint8_t*)w->out.p + (b_lo - first * h.blocksize), 1, (size_t)(b_hi - b_lo), w->stream)) break;
    result = (int)(b_hi - b_lo);
  } while (0);
  ws_release(w);
  return result;
}

int blosc_getitem(const void* src, int start, int nitems, void* dest) { return getitem_impl(src, start, nitems, dest, -1); }

/* ------------------------------------------------------------------------- */
/* frames: buffers larger than one chunk           */
/* ------------------------------------------------------------------------- */
/* A Blosc-1 chunk holds at most INT_MAX-16 bytes (blosc.h:40) and a single call keeps the GPU
 * about 7 % busy, so large buffers are cut into independent chunks that
 * are compressed by a few host threads at once, each on its own workspace and stream: chunk
 * i+1's PCIe transfer and filter overlap chunk i's codec kernels.  The container is a minimal
 * index in front of ordinary chunks, every one of which the reference library can decode:
 *
 *   0   "B2FR"            magic
 *   4   u8 version (1), 3 reserved bytes (0)
 *   8   u64 nbytes        uncompressed size of the whole frame
 *   16  u64 cbytes        size of the frame itself
 *   24  u32 chunksize     uncompressed bytes per chunk (the last one may be shorter)
 *   28  u32 nchunks
 *   32  u64 offset[nchunks]   from the start of the frame
 *   ..  the chunks, back to back, in order
 */
#define B2_FRAME_HDR 32
#define B2_FRAME_DEFAULT_CHUNK ((size_t)256 << 20)
#define B2_FRAME_MAX_WORKERS 8

struct b2_frame_job {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int next, commit, failed, err, nchunks, dev, dest_dev, workers;
  /* compression */
  int clevel, doshuffle, nthreads;
  size_t typesize, blocksize, chunksize, nbytes, destsize, cursor;
  const char* compressor;
  const uint8_t* src;
  uint8_t* dest;
  uint64_t* offsets;
  /* decompression */
  const uint8_t* frame;
};

static void wr_u64(uint8_t* p, uint64_t v) { int i; for (i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static uint64_t rd_u64(const uint8_t* p) { uint64_t v = 0; int i; for (i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
static void wr_u32(uint8_t* p, uint32_t v) { wr_i32(p, (int32_t)v); }
static uint32_t rd_u32(const uint8_t* p) { return (uint32_t)rd_i32(p); }

static int frame_workers(int nchunks) {
  const char* e = getenv("BLOSC_B200_FRAME_WORKERS");
  int n = e ? atoi(e) : 4;
  if (n < 1) n = 1;
  if (n > B2_FRAME_MAX_WORKERS) n = B2_FRAME_MAX_WORKERS;
  return n < nchunks ? n : nchunks;
}

/* Ordered commit: chunk `index` gets the bytes right after chunk index-1.  Returns NULL (and
 * marks the job failed) when the frame does not fit; cbytes < 0 gives up the turn after an error. */
static void* frame_place(b2_place* pl, int32_t cbytes) {
  b2_frame_job* j = pl->job;
  void* where = NULL;
  pthread_mutex_lock(&j->mu);
  while (j->commit != pl->index) pthread_cond_wait(&j->cv, &j->mu);
  if (cbytes >= 0 && !j->failed && j->cursor + (size_t)cbytes <= j->destsize) {
    j->offsets[pl->index] = j->cursor;
    where = j->dest + j->cursor;
    j->cursor += (size_t)cbytes;
  } else j->failed = 1;
  pl->placed = 1;
  j->commit++;
  pthread_cond_broadcast(&j->cv);
  pthread_mutex_unlock(&j->mu);
  return where;
}

static int frame_take(b2_frame_job* j, int* failed) {
  int i;
  pthread_mutex_lock(&j->mu);
  i = j->next < j->nchunks ? j->next++ : -1;
  *failed = j->failed;
  pthread_mutex_unlock(&j->mu);
  return i;
}

static void frame_fail(b2_frame_job* j, int rc) {
  pthread_mutex_lock(&j->mu);
  j->failed = 1;
  if (rc < 0 && !j->err) j->err = rc;
  pthread_mutex_unlock(&j->mu);
}

static void* frame_compress_worker(void* arg) {
  b2_frame_job* j = (b2_frame_job*)arg;
  int i, failed;
  b2_set_device(j->dev);
  while ((i = frame_take(j, &failed)) >= 0) {
    b2_place pl;
    const size_t off = (size_t)i * j->chunksize;
    const size_t n = j->nbytes - off < j->chunksize ? j->nbytes - off : j->chunksize;
    int rc = -1;
    pl.job = j; pl.index = i; pl.placed = 0; pl.many = j->workers > 1;
    if (!failed)
      rc = compress_impl(j->clevel, j->doshuffle, j->typesize, n, j->src + off, NULL, n + BLOSC_MAX_OVERHEAD,
                         j->compressor, j->blocksize, j->nthreads, &pl, j->dest_dev);
    if (!pl.placed) frame_place(&pl, -1);
    if (rc <= 0 && !failed) frame_fail(j, rc);
  }
  return NULL;
}

static void* frame_decompress_worker(void* arg) {
  b2_frame_job* j = (b2_frame_job*)arg;
  int i, failed;
  b2_set_device(j->dev);
  while ((i = frame_take(j, &failed)) >= 0) {
    const size_t off = (size_t)i * j->chunksize;
    const size_t n = j->nbytes - off < j->chunksize ? j->nbytes - off : j->chunksize;
    int rc;
    if (failed) continue;
    rc = decompress_impl(j->frame + j->offsets[i], j->dest + off, n, j->nthreads,
                         (long long)(j->offsets[i + 1] - j->offsets[i]), (long long)n);
    if (rc != (int)n) frame_fail(j, -1);
  }
  return NULL;
}

static int frame_run(b2_frame_job* j, void* (*fn)(void*)) {
  pthread_t th[B2_FRAME_MAX_WORKERS];
  int k, started = 0;
  pthread_mutex_init(&j->mu, NULL);
  pthread_cond_init(&j->cv, NULL);
  j->dev = b2_get_device();
  for (k = 1; k < j->workers; k++) {
    if (pthread_create(&th[started], NULL, fn, j) != 0) break;
    started++;
  }
  fn(j);                                   /* the calling thread is worker 0 */
  for (k = 0; k < started; k++) pthread_join(th[k], NULL);
  pthread_cond_destroy(&j->cv);
  pthread_mutex_destroy(&j->mu);
  return j->failed ? -1 : 0;
}

static size_t frame_chunksize(size_t chunksize, size_t typesize) {
  if (chunksize == 0) chunksize = B2_FRAME_DEFAULT_CHUNK;
  if (chunksize > BLOSC_MAX_BUFFERSIZE) chunksize = BLOSC_MAX_BUFFERSIZE;
  if (typesize > 1 && typesize <= BLOSC_MAX_TYPESIZE && chunksize >= typesize) chunksize -= chunksize % typesize;
  return chunksize;
}

size_t blosc_b200_frame_bound(size_t nbytes, size_t typesize, size_t chunksize) {
  size_t nchunks;
  chunksize = frame_chunksize(chunksize, typesize);
  nchunks = (nbytes + chunksize - 1) / chunksize;
  return B2_FRAME_HDR + nchunks * (8 + BLOSC_MAX_OVERHEAD) + nbytes;
}

long long blosc_b200_frame_compress(int clevel, int doshuffle, size_t typesize, size_t nbytes, const void* src,
                                    void* dest, size_t destsize, const char* compressor, size_t blocksize,
                                    size_t chunksize, int numinternalthreads) {
  b2_frame_job j;
  uint8_t* index;
  size_t nchunks, index_bytes;
  int i, rc, dest_dev;

  if (clevel < 0 || clevel > 9) return -10;                       /* same codes as the chunk API */
  if (doshuffle != 0 && doshuffle != 1 && doshuffle != 2) return -10;
  if (typesize == 0) return -10;
  if (blosc_compname_to_compcode(compressor) < 0) return -5;
  chunksize = frame_chunksize(chunksize, typesize);
  nchunks = (nbytes + chunksize - 1) / chunksize;
  if (nchunks > 0x7fffffff / 2) return -1;
  index_bytes = B2_FRAME_HDR + nchunks * 8;
  if (destsize < index_bytes) return 0;
  if (!backend_ready()) return -1;
  dest_dev = b2_ptr_is_device(dest);

  memset(&j, 0, sizeof j);
  j.nchunks = (int)nchunks; j.workers = frame_workers((int)nchunks); j.dest_dev = dest_dev;
  j.clevel = clevel; j.doshuffle = doshuffle; j.nthreads = numinternalthreads;
  j.typesize = typesize; j.blocksize = blocksize; j.chunksize = chunksize; j.nbytes = nbytes;
  j.destsize = destsize; j.cursor = index_bytes; j.compressor = compressor;
  j.src = (const uint8_t*)src; j.dest = (uint8_t*)dest;
  j.offsets = (uint64_t*)calloc(nchunks ? nchunks : 1, sizeof(uint64_t));
  index = (uint8_t*)malloc(index_bytes);
  if (!j.offsets || !index) { free(j.offsets); free(index); return -1; }
  rc = nchunks ? frame_run(&j, frame_compress_worker) : 0;
  if (rc == 0) {
    memcpy(index, "B2FR", 4); index[4] = 1; index[5] = index[6] = index[7] = 0;
    wr_u64(index + 8, (uint64_t)nbytes); wr_u64(index + 16, (uint64_t)j.cursor);
    wr_u32(index + 24, (uint32_t)chunksize); wr_u32(index + 28, (uint32_t)nchunks);
    for (i = 0; i < (int)nchunks; i++) wr_u64(index + B2_FRAME_HDR + 8 * (size_t)i, j.offsets[i]);
    if (dest_dev) {
      b2_ws* w = ws_acquire();
      rc = w ? copy_any(dest, 1, index, 0, index_bytes, w->stream) : -1;
      if (w) ws_release(w);
    } else memcpy(dest, index, index_bytes);
  }
  free(j.offsets); free(index);
  if (rc) return j.err ? j.err : (j.failed ? 0 : -1);            /* 0: does not fit in destsize, as blosc_compress */
  return (long long)j.cursor;
}

/* reads and validates the index; *offsets is malloc'ed (nchunks+1 entries, the last one = cbytes) */
static int frame_open(const void* frame, size_t framesize, size_t* nbytes, size_t* chunksize, size_t* nchunks,
                      uint64_t** offsets) {
  uint8_t hb[B2_FRAME_HDR];
  uint8_t* raw;
  uint64_t* off;
  uint64_t cbytes;
  size_t n, k;
  const int dev = b2_ptr_is_device(frame);
  b2_ws* w = NULL;
  int rc = 0;
  if (framesize < B2_FRAME_HDR) return -1;
  if (dev) {
    w = ws_acquire();
    if (!w) return -1;
    rc = copy_any(hb, 0, frame, 1, B2_FRAME_HDR, w->stream);
    if (rc) { ws_release(w); return -1; }
  } else memcpy(hb, frame, B2_FRAME_HDR);
  rc = -1;
  do {
    if (memcmp(hb, "B2FR", 4) != 0 || hb[4] != 1) break;
    *nbytes = (size_t)rd_u64(hb + 8); cbytes = rd_u64(hb + 16);
    *chunksize = rd_u32(hb + 24); n = rd_u32(hb + 28);
    if (cbytes > framesize || cbytes < B2_FRAME_HDR + 8 * (uint64_t)n) break;
    if (*nbytes > 0 && (*chunksize == 0 || *chunksize > BLOSC_MAX_BUFFERSIZE)) break;
    if (n != (*nbytes ? (*nbytes + *chunksize - 1) / *chunksize : 0)) break;
    raw = (uint8_t*)malloc(8 * n + 8);
    off = (uint64_t*)malloc(8 * (n + 1));
    if (!raw || !off) { free(raw); free(off); break; }
    if (dev) { if (copy_any(raw, 0, (const uint8_t*)frame + B2_FRAME_HDR, 1, 8 * n, w->stream)) { free(raw); free(off); break; } }
    else memcpy(raw, (const uint8_t*)frame + B2_FRAME_HDR, 8 * n);
    for (k = 0; k < n; k++) off[k] = rd_u64(raw + 8 * k);
    off[n] = cbytes;
    free(raw);
    rc = 0;
    for (k = 0; k < n; k++)                /* chunks in order, at least a header each, inside the frame */
      if (off[k] < B2_FRAME_HDR + 8 * (uint64_t)n || off[k + 1] < off[k] + BLOSC_MAX_OVERHEAD || off[k + 1] > cbytes) rc = -1;
    if (rc) { free(off); break; }
    *nchunks = n; *offsets = off;
  } while (0);
  if (w) ws_release(w);
  return rc;
}

int blosc_b200_frame_info(const void* frame, size_t framesize, size_t* nbytes, size_t* cbytes, size_t* chunksize,
                          size_t* nchunks) {
  size_t nb = 0, cs = 0, nc = 0;
  uint64_t* off = NULL;
  if (!backend_ready()) return -1;
  if (frame_open(frame, framesize, &nb, &cs, &nc, &off)) return -1;
  if (nbytes) *nbytes = nb;
  if (cbytes) *cbytes = (size_t)off[nc];
  if (chunksize) *chunksize = cs;
  if (nchunks) *nchunks = nc;
  free(off);
  return 0;
}

long long blosc_b200_frame_chunk(const void* frame, size_t framesize, size_t i, size_t* chunk_cbytes) {
  size_t nb = 0, cs = 0, nc = 0;
  uint64_t* off = NULL;
  long long r;
  if (!backend_ready()) return -1;
  if (frame_open(frame, framesize, &nb, &cs, &nc, &off)) return -1;
  if (i >= nc) { free(off); return -1; }
  if (chunk_cbytes) *chunk_cbytes = (size_t)(off[i + 1] - off[i]);
  r = (long long)off[i];
  free(off);
  return r;
}

long long blosc_b200_frame_decompress(const void* frame, size_t framesize, void* dest, size_t destsize,
                                      int numinternalthreads) {
  b2_frame_job j;
  size_t nb = 0, cs = 0, nc = 0;
  uint64_t* off = NULL;
  int rc;
  if (!backend_ready()) return -1;
  if (frame_open(frame, framesize, &nb, &cs, &nc, &off)) return -1;
  if (nb > destsize) { free(off); return -1; }
  memset(&j, 0, sizeof j);
  j.nchunks = (int)nc; j.workers = frame_workers((int)nc); j.nthreads = numinternalthreads;
  j.chunksize = cs; j.nbytes = nb; j.frame = (const uint8_t*)frame; j.dest = (uint8_t*)dest; j.offsets = off;
  rc = nc ? frame_run(&j, frame_decompress_worker) : 0;
  free(off);
  return rc ? -1 : (long long)nb;
}

long long blosc_b200_frame_getitem(const void* frame, size_t framesize, size_t start, size_t nitems, void* dest) {
  size_t nb = 0, cs = 0, nc = 0, ts = 0, ipc, done = 0;
  uint64_t* off = NULL;
  uint8_t hb[16];
  long long result = -1;
  if (!backend_ready()) return -1;
  if (frame_open(frame, framesize, &nb, &cs, &nc, &off)) return -1;
  do {
    if (nc == 0) { result = nitems == 0 ? 0 : -1; break; }
    if (b2_ptr_is_device(frame)) {
      b2_ws* w = ws_acquire();
      int rc = w ? copy_any(hb, 0, (const uint8_t*)frame + off[0], 1, 16, w->stream) : -1;
      if (w) ws_release(w);
      if (rc) break;
    } else memcpy(hb, (const uint8_t*)frame + off[0], 16);
    ts = hb[3];
    if (ts == 0 || cs % ts) break;
    ipc = cs / ts;                                     /* items per chunk */
    if (start > nb / ts || nitems > nb / ts - start) { fprintf(stderr, "`start`+`nitems` out of bounds"); break; }
    result = 0;
    while (done < nitems) {
      const size_t c = (start + done) / ipc, first = (start + done) % ipc;
      const size_t take = nitems - done < ipc - first ? nitems - done : ipc - first;
      const int rc = getitem_impl((const uint8_t*)frame + off[c], (int)first, (int)take, (uint8_t*)dest + done * ts,
                                  (long long)(off[c + 1] - off[c]));
      if (rc != (int)(take * ts)) { result = rc < 0 ? rc : -1; break; }
      done += take;
      result += rc;
    }
  } while (0);
  free(off);
  return result;
}

/* ------------------------------------------------------------------------- */
/* global-state front end                                                     */
/* ------------------------------------------------------------------------- */
void blosc_init(void) { g_initlib = 1; }                                    /* blosc.c:2223-2247 (no pool to create) */
void blosc_destroy(void) { if (g_initlib) { g_initlib = 0; blosc_free_resources(); } }   /* blosc.c:2249-2260 */
int blosc_get_nthreads(void) { return g_threads; }
int blosc_set_nthreads(int n) { int old = g_threads; if (!g_initlib) blosc_init(); g_threads = n; return old; }   /* blosc.c:1958-1975 */
const char* blosc_get_compressor(void) { const char* n; blosc_compcode_to_compname(g_compressor, &n); return n; }
int blosc_set_compressor(const char* compname) {                            /* blosc.c:2013-2023 */
  int code = blosc_compname_to_compcode(compname);
  g_compressor = code;
  if (!g_initlib) blosc_init();
  return code;
}
int blosc_get_blocksize(void) { return g_force_blocksize; }
void blosc_set_blocksize(size_t size) { g_force_blocksize = (int32_t)size; }
void blosc_set_splitmode(int mode) { g_splitmode = mode; }

int blosc_compress(int clevel, int doshuffle, size_t typesize, size_t nbytes, const void* src, void* dest,
                   size_t destsize) {                                      /* blosc.c:1311-1433 */
  const char* envvar;
  const char* compname;
  int result, nolock;
  if (!g_initlib) blosc_init();
  if ((envvar = getenv("BLOSC_CLEVEL")) != NULL) { long v = strtol(envvar, NULL, 10); if (v != EINVAL && v >= 0) clevel = (int)v; }
  if ((envvar = getenv("BLOSC_SHUFFLE")) != NULL) {
    if (strcmp(envvar, "NOSHUFFLE") == 0) doshuffle = BLOSC_NOSHUFFLE;
    if (strcmp(envvar, "SHUFFLE") == 0) doshuffle = BLOSC_SHUFFLE;
    if (strcmp(envvar, "BITSHUFFLE") == 0) doshuffle = BLOSC_BITSHUFFLE;
  }
  if ((envvar = getenv("BLOSC_TYPESIZE")) != NULL) { long v = strtol(envvar, NULL, 10); if (v != EINVAL && v > 0) typesize = (size_t)(int)v; }
  if ((envvar = getenv("BLOSC_COMPRESSOR")) != NULL) { result = blosc_set_compressor(envvar); if (result < 0) return result; }
  if ((envvar = getenv("BLOSC_BLOCKSIZE")) != NULL) { long v = strtol(envvar, NULL, 10); if (v != EINVAL && v > 0) blosc_set_blocksize((size_t)v); }
  if ((envvar = getenv("BLOSC_NTHREADS")) != NULL) { long v = strtol(envvar, NULL, 10); if (v != EINVAL && v > 0) { result = blosc_set_nthreads((int)v); if (result < 0) return result; } }
  if ((envvar = getenv("BLOSC_SPLITMODE")) != NULL) {
    if (strcmp(envvar, "FORWARD_COMPAT") == 0) blosc_set_splitmode(BLOSC_FORWARD_COMPAT_SPLIT);
    else if (strcmp(envvar, "AUTO") == 0) blosc_set_splitmode(BLOSC_AUTO_SPLIT);
    else if (strcmp(envvar, "ALWAYS") == 0) blosc_set_splitmode(BLOSC_ALWAYS_SPLIT);
    else if (strcmp(envvar, "NEVER") == 0) blosc_set_splitmode(BLOSC_NEVER_SPLIT);
    else { fprintf(stderr, "BLOSC_SPLITMODE environment variable '%s' not recognized\n", envvar); return -1; }
  }
  nolock = getenv("BLOSC_NOLOCK") != NULL;
  blosc_compcode_to_compname(g_compressor, &compname);
  if (compname == NULL) compname = "(null)";
  /* the global path serialises callers on one mutex (blosc.c:1410); BLOSC_NOLOCK skips it (:1400-1408) */
  if (!nolock) pthread_mutex_lock(&g_global_mutex);
  result = blosc_compress_ctx(clevel, doshuffle, typesize, nbytes, src, dest, destsize, compname,
                              (size_t)g_force_blocksize, g_threads);
  if (!nolock) pthread_mutex_unlock(&g_global_mutex);
  return result;
}

int blosc_decompress(const void* src, void* dest, size_t destsize) {        /* blosc.c:1537-1572 */
  const char* envvar;
  int result, nolock;
  if (!g_initlib) blosc_init();
  if ((envvar = getenv("BLOSC_NTHREADS")) != NULL) { long v = strtol(envvar, NULL, 10); if (v != EINVAL && v > 0) { result = blosc_set_nthreads((int)v); if (result < 0) return result; } }
  nolock = getenv("BLOSC_NOLOCK") != NULL;
  if (!nolock) pthread_mutex_lock(&g_global_mutex);
  result = blosc_decompress_ctx(src, dest, destsize, g_threads);
  if (!nolock) pthread_mutex_unlock(&g_global_mutex);
  return result;
}

/* ------------------------------------------------------------------------- */
/* B200 extensions                                                            */
/* ------------------------------------------------------------------------- */
int blosc_b200_filter(int mode, size_t typesize, size_t blocksize, const void* src, void* dest) {
  b2_ws* w;
  FilterArgs fa;
  int src_dev, dest_dev, rc = -1;
  if (mode < 0 || mode > 3 || typesize == 0 || blocksize > (size_t)INT_MAX) return -1;
  if (blocksize == 0) return 0;
  src_dev = b2_ptr_is_device(src); dest_dev = b2_ptr_is_device(dest);
  w = ws_acquire();
  if (!w) return -1;
  do {
    const uint8_t* d_src = (const uint8_t*)src;
    uint8_t* d_dst = (uint8_t*)dest;
    if (!src_dev) {
      if (buf_ensure(&w->in, blocksize + 64)) break;
      if (b2_copy_h2d(w->in.p, src, blocksize, w->stream)) break;
      d_src = (const uint8_t*)w->in.p;
    }
    if (!dest_dev) { if (buf_ensure(&w->out, blocksize + 64)) break; d_dst = (uint8_t*)w->out.p; }
    fa.src = d_src; fa.dst = d_dst; fa.nbytes = (long long)blocksize; fa.blocksize = (int)blocksize;
    fa.typesize = (int)typesize; fa.mode = mode;
    if (b2_launch_filter(&fa, w->stream)) break;
    if (!dest_dev && b2_copy_d2h(dest, d_dst, blocksize, w->stream)) break;
    if (b2_stream_sync(w->stream)) break;
    rc = 0;
  } while (0);
  ws_release(w);
  return rc;
}

int blosc_b200_set_device(int dev) { return backend_ready() ? b2_set_device(dev) : -1; }
void blosc_b200_set_profiling(int on) { b2_prof_enable(on); }
void blosc_b200_prof_reset(void) { b2_prof_reset(); }
int blosc_b200_prof_get(int kind, double* ms_total, long long* launches) { return b2_prof_get(kind, ms_total, launches); }
long long blosc_b200_launch_count(void) { return b2_launch_count(); }
