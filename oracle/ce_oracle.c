/* oracle/ce_oracle.c -- CPU restatement of the reference's hot path, in plain C99.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (catears_b200 / libce_gpu.so) never does and has no CPU fallback.
 *
 * Every function restates one reference function in its own words and cites
 * it (paths are relative to the reference root).  The restatement is PINNED:
 * tests/test_oracle.py checks it against the reference's own known answers
 * (test/srfft_test.cc 128-point vector, the Kaldi fbank / online-CMVN dumps in
 * test/data, the literals of test/nnet_test.cc) and, where /root/reference was
 * available at build time, against the reference itself compiled unmodified
 * (oracle/_ref/libce_ref.so).
 *
 * Build with -ffp-contract=off: the reference is compiled for baseline x86-64
 * (Makefile.am:4, no -mfma), so every fp32 multiply and add rounds separately.
 *
 * Deliberate difference: the N/2-point complex FFT is an iterative radix-2
 * (twiddles rounded from double) instead of the reference's recursive
 * split-radix (src/srfft.cc:124-265); both compute the same un-normalised DFT
 * and agree to fp32 rounding.  The real-FFT post-pass keeps the reference's
 * fp32 twiddle recurrence (src/srfft.cc:383-392).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SAMPLE_RATE 16000      /* src/fbank.h:7  */
#define ORC_FRAME_SHIFT 160        /* src/fbank.h:8, src/fbank.cc:15 */
#define ORC_FRAME_LEN 400          /* src/fbank.h:9, src/fbank.cc:16 */
#define ORC_PADDED 512             /* src/fbank.cc:258 */
#define ORC_LOW_FREQ 20            /* src/fbank.h:11 */
#define ORC_HIGH_FREQ 8000         /* src/fbank.h:12 */
#define ORC_PREEMPH 0.97           /* src/fbank.h:13 (a double literal) */
#define ORC_CMVN_WINDOW 600        /* src/cmvn.h:10 */
#define ORC_CMVN_GLOBAL 200        /* src/cmvn.h:11 */
#define ORC_MAX_MEL 128

/* ------------------------------------------------------------------------- */
/* FFT                                                                        */
/* ------------------------------------------------------------------------- */

/* In-place forward complex FFT of n points (re/im interleaved), iterative
 * radix-2 decimation in time.  Contract of src/srfft.cc:293-340. */
static void complex_fft(float *x, int n) {
  int i, j, len;
  for (i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      float tr = x[2 * i], ti = x[2 * i + 1];
      x[2 * i] = x[2 * j]; x[2 * i + 1] = x[2 * j + 1];
      x[2 * j] = tr; x[2 * j + 1] = ti;
    }
  }
  for (len = 2; len <= n; len <<= 1) {
    int half = len >> 1;
    for (j = 0; j < half; ++j) {
      double ang = -2.0 * 3.14159265358979323846 * (double)j / (double)len;
      float wr = (float)cos(ang), wi = (float)sin(ang);
      for (i = j; i < n; i += len) {
        float *a = x + 2 * i, *b = x + 2 * (i + half);
        float tr = b[0] * wr - b[1] * wi;
        float ti = b[0] * wi + b[1] * wr;
        b[0] = a[0] - tr; b[1] = a[1] - ti;
        a[0] = a[0] + tr; a[1] = a[1] + ti;
      }
    }
  }
}

/* Forward real FFT of n floats, output packed [Re0, Re(n/2), Re1, Im1, ...].
 * src/srfft.cc:370-445 (forward branch). */
void orc_srfft(float *data, int n) {
  int N = n, N2 = n / 2, k;
  float root_re, root_im, kn_re = 1.0f, kn_im = 0.0f;
  complex_fft(data, N2);
  {
    /* src/srfft.cc:385: the angle is rounded to float, then cos/sin of a float. */
    float ang = (float)(6.283185307179586476925286766559005 / N * -1);
    root_re = cosf(ang);
    root_im = sinf(ang);
  }
  for (k = 1; 2 * k <= N2; ++k) {
    float ck_re, ck_im, dk_re, dk_im, t;
    int kd = N2 - k;
    /* kN *= rootN in fp32 (src/srfft.cc:392, complex_mul :53-57). */
    t = kn_re * root_re - kn_im * root_im;
    kn_im = kn_re * root_im + kn_im * root_re;
    kn_re = t;
    ck_re = 0.5f * (data[2 * k] + data[N - 2 * k]);
    ck_im = 0.5f * (data[2 * k + 1] - data[N - 2 * k + 1]);
    dk_re = 0.5f * (data[2 * k + 1] + data[N - 2 * k + 1]);
    dk_im = -0.5f * (data[2 * k] - data[N - 2 * k]);
    /* A_k = C_k + kN * D_k (complex_add_product :59-68). */
    data[2 * k] = ck_re + (kn_re * dk_re - kn_im * dk_im);
    data[2 * k + 1] = ck_im + (kn_re * dk_im + kn_im * dk_re);
    if (kd != k) {
      /* conj(C_k), conj(D_k), twiddle (-kn_re, kn_im)  (:414-433) */
      float ndk_im = -dk_im, nkn_re = -kn_re;
      data[2 * kd] = ck_re + (nkn_re * dk_re - kn_im * ndk_im);
      data[2 * kd + 1] = -ck_im + (nkn_re * ndk_im + kn_im * dk_re);
    }
  }
  {
    float zeroth = data[0] + data[1], n2th = data[0] - data[1];  /* :442-445 */
    data[0] = zeroth;
    data[1] = n2th;
  }
}

/* ------------------------------------------------------------------------- */
/* Fbank                                                                      */
/* ------------------------------------------------------------------------- */

typedef struct {
  int mel;
  float hamming[ORC_FRAME_LEN];
  int offset[ORC_MAX_MEL];
  int width[ORC_MAX_MEL];
  float *weights[ORC_MAX_MEL];
} orc_fbank_t;

static float mel_scale(float freq) { /* src/fbank.h:30-32 */
  return 1127.0f * logf(1.0f + freq / 700.0f);
}

/* Hamming table (src/fbank.cc:248-255) and triangular mel filters
 * (src/fbank.cc:103-163).  All arithmetic in fp32 as there. */
/* strict != 0 keeps the reference's assertion that every filter spans >= 2 FFT bins
 * (fbank.cc:155); with 80 bins the lowest filter holds a single bin, so the reference
 * itself aborts there and only strict == 0 can build the tables. */
orc_fbank_t *orc_fbank_create2(int mel, int strict) {
  int i, b;
  orc_fbank_t *fb;
  if (mel < 3 || mel > ORC_MAX_MEL) return NULL;
  fb = (orc_fbank_t *)calloc(1, sizeof(orc_fbank_t));
  fb->mel = mel;
  {
    float a = (float)(6.28318530718 / (ORC_FRAME_LEN - 1));   /* fbank.cc:19,250 */
    for (i = 0; i < ORC_FRAME_LEN; ++i) {
      float fi = (float)i;
      fb->hamming[i] = (float)(0.54 - 0.46 * cosf(a * fi));
    }
  }
  {
    int num_fft_bins = ORC_PADDED / 2;
    float sample_freq = ORC_SAMPLE_RATE;
    float bin_width = sample_freq / ORC_PADDED;
    float mel_low = mel_scale(ORC_LOW_FREQ), mel_high = mel_scale(ORC_HIGH_FREQ);
    float delta = (mel_high - mel_low) / (mel + 1);
    float tmp[ORC_PADDED / 2];
    for (b = 0; b < mel; ++b) {
      float left = mel_low + b * delta;
      float center = mel_low + (b + 1) * delta;
      float right = mel_low + (b + 2) * delta;
      int first = -1, last = -1;
      for (i = 0; i < num_fft_bins; ++i) {
        float m = mel_scale(bin_width * i);
        tmp[i] = 0.0f;
        if (m > left && m < right) {               /* strict, fbank.cc:141 */
          tmp[i] = (m <= center) ? (m - left) / (center - left)
                                 : (right - m) / (right - center);
          if (first == -1) first = i;
          last = i;
        }
      }
      if (first == -1 || (strict && last <= first)) { free(fb); return NULL; }  /* assert fbank.cc:155 */
      fb->offset[b] = first;
      fb->width[b] = last + 1 - first;
      fb->weights[b] = (float *)malloc(sizeof(float) * fb->width[b]);
      memcpy(fb->weights[b], tmp + first, sizeof(float) * fb->width[b]);
    }
  }
  return fb;
}

orc_fbank_t *orc_fbank_create(int mel) { return orc_fbank_create2(mel, 1); }

void orc_fbank_destroy(orc_fbank_t *fb) {
  int b;
  if (!fb) return;
  for (b = 0; b < fb->mel; ++b) free(fb->weights[b]);
  free(fb);
}

/* Table access for the tests (mel filter k: offset, width, weights). */
int orc_fbank_filter(const orc_fbank_t *fb, int b, int *offset, float *w, int cap) {
  if (b < 0 || b >= fb->mel || fb->width[b] > cap) return -1;
  *offset = fb->offset[b];
  memcpy(w, fb->weights[b], sizeof(float) * fb->width[b]);
  return fb->width[b];
}
const float *orc_fbank_hamming(const orc_fbank_t *fb) { return fb->hamming; }

int orc_num_frames(int n_samples) { /* src/fbank.cc:35-42 */
  return n_samples < ORC_FRAME_LEN ? 0 : 1 + (n_samples - ORC_FRAME_LEN) / ORC_FRAME_SHIFT;
}

/* One frame: src/fbank.cc:74-100 (ExtractWindow), :44-69 (ProcessWindow),
 * :219-245 (ComputeFrame), :193-211 (ComputePowerSpectrum), :165-184 (Melbanks::Compute). */
static void fbank_frame(const orc_fbank_t *fb, const float *wave, float *out) {
  float win[ORC_PADDED];
  float sum = 0.0f, mean;
  int i, b;
  for (i = 0; i < ORC_FRAME_LEN; ++i) win[i] = wave[i];
  for (i = ORC_FRAME_LEN; i < ORC_PADDED; ++i) win[i] = 0.0f;
  for (i = 0; i < ORC_FRAME_LEN; ++i) sum += win[i];
  mean = sum / ORC_FRAME_LEN;
  for (i = 0; i < ORC_FRAME_LEN; ++i) win[i] -= mean;
  for (i = ORC_FRAME_LEN - 1; i > 0; --i)                /* double multiply-subtract */
    win[i] = (float)((double)win[i] - ORC_PREEMPH * (double)win[i - 1]);
  win[0] = (float)((double)win[0] - ORC_PREEMPH * (double)win[0]);   /* fbank.cc:61 */
  for (i = 0; i < ORC_FRAME_LEN; ++i) win[i] *= fb->hamming[i];

  orc_srfft(win, ORC_PADDED);
  {
    int half = ORC_PADDED / 2;
    float first = win[0] * win[0], last = win[1] * win[1];
    for (i = 1; i < half; ++i) {
      float re = win[2 * i], im = win[2 * i + 1];
      win[i] = re * re + im * im;
    }
    win[0] = first;
    win[half] = last;
  }
  for (b = 0; b < fb->mel; ++b) {
    const float *w = fb->weights[b];
    const float *p = win + fb->offset[b];
    float e = 0.0f;
    for (i = 0; i < fb->width[b]; ++i) e += w[i] * p[i];  /* VecVec, vector.cc:82-92 */
    if (e < FLT_EPSILON) e = FLT_EPSILON;                 /* fbank.cc:243 */
    out[b] = logf(e);                                     /* fbank.cc:244 */
  }
}

/* Whole-buffer fbank of int16 PCM, unscaled (src/pcm_reader.cc:168-182,
 * src/fbank.cc:265-303).  Returns the frame count; out is [frames x mel]. */
int orc_fbank(const orc_fbank_t *fb, const int16_t *pcm, int n_samples, float *out) {
  int T = orc_num_frames(n_samples), t, i;
  float wave[ORC_FRAME_LEN];
  for (t = 0; t < T; ++t) {
    const int16_t *p = pcm + (size_t)t * ORC_FRAME_SHIFT;
    for (i = 0; i < ORC_FRAME_LEN; ++i) wave[i] = (float)p[i];
    fbank_frame(fb, wave, out + (size_t)t * fb->mel);
  }
  return T;
}

/* ------------------------------------------------------------------------- */
/* Online CMVN (mean only)                                                    */
/* ------------------------------------------------------------------------- */

/* src/cmvn.cc:100-110 = ComputeStats :35-68 + SmoothStats :70-89 + Apply :91-98,
 * for frames 0..T-1 in order.  `g` = global stats [mel sums, count]. */
void orc_cmvn(const float *g, const float *feats, int T, int mel, float *out) {
  float cached[ORC_MAX_MEL + 1];
  float stats[ORC_MAX_MEL + 1];
  int t, d;
  for (d = 0; d <= mel; ++d) cached[d] = 0.0f;
  for (t = 0; t < T; ++t) {
    const float *x = feats + (size_t)t * mel;
    double count;
    for (d = 0; d < mel; ++d) {
      double s = (double)cached[d];
      s += (double)x[d];
      if (t - ORC_CMVN_WINDOW >= 0)
        s += -1.0 * (double)feats[(size_t)(t - ORC_CMVN_WINDOW) * mel + d];
      cached[d] = (float)s;
    }
    {
      double c = (double)cached[mel] + 1.0;
      if (t - ORC_CMVN_WINDOW >= 0) c -= 1.0;
      cached[mel] = (float)c;
    }
    for (d = 0; d <= mel; ++d) stats[d] = cached[d];
    count = (double)stats[mel];
    if (count < ORC_CMVN_WINDOW) {                       /* SmoothStats */
      double from_global = ORC_CMVN_WINDOW - count;
      double global_count = (double)g[mel];
      float alpha;
      if (from_global > ORC_CMVN_GLOBAL) from_global = ORC_CMVN_GLOBAL;
      alpha = (float)(from_global / global_count);       /* AddVec takes a float alpha */
      for (d = 0; d <= mel; ++d) stats[d] += alpha * g[d];
    }
    {
      double cnt = (double)stats[mel];
      float scale = (float)(1 / cnt);                    /* cmvn.cc:96 */
      float nscale = -scale;
      for (d = 0; d < mel; ++d) {
        float y = x[d];
        y += nscale * stats[d];                          /* AddVec(-scale, stats) */
        out[(size_t)t * mel + d] = y;
      }
    }
  }
}

/* ------------------------------------------------------------------------- */
/* Matrix / quantisation                                                      */
/* ------------------------------------------------------------------------- */

/* C = A*B, fp32, k ascending per element: the summation order of SimpleMatMat
 * (src/matrix.cc:275-292); stands in for cblas_sgemm (src/matrix.cc:300-323). */
void orc_sgemm(const float *a, const float *b, int m, int n, int k, float *c) {
  int i, j, kk;
  for (i = 0; i < m; ++i) {
    float *crow = c + (size_t)i * n;
    for (j = 0; j < n; ++j) crow[j] = 0.0f;
    for (kk = 0; kk < k; ++kk) {
      float av = a[(size_t)i * k + kk];
      const float *brow = b + (size_t)kk * n;
      for (j = 0; j < n; ++j) crow[j] += av * brow[j];
    }
  }
}

/* src/matrix.cc:329-362 (FindMinMax with max seeded by FLT_MIN, the smallest
 * positive normal; ComputeQuantizationParams) and :366-387 (Quantize). */
void orc_quantize(const float *src, long count, uint8_t *dst, float *scale_out,
                  int32_t *zp_out) {
  float mn = FLT_MAX, mx = FLT_MIN;
  long i;
  double scale, fzp;
  int32_t zp;
  float scale_f;
  for (i = 0; i < count; ++i) {
    float v = src[i];
    if (v > mx) mx = v;
    if (v < mn) mn = v;
  }
  scale = (mx - mn) / 255.0;
  fzp = -mn / scale;
  zp = (int32_t)round(fzp);
  scale_f = (float)scale;
  for (i = 0; i < count; ++i) {
    float v = src[i];
    v = v / scale_f + zp;
    if (v > 255.0f) v = 255.0f;
    if (v < 0.0f) v = 0.0f;
    dst[i] = (uint8_t)roundf(v);
  }
  *scale_out = scale_f;
  *zp_out = zp;
}

/* acc[i][j] = sum_k (A[i,k]-zpA)(B[k,j]-zpB) in int32 -- the arithmetic contract
 * of gemmlowp with the empty output pipeline (eight_bit_int_gemm.cc:107-133,
 * internal/unpack.h:118-125); C = (float)acc * (sA*sB) (eight_bit_int_gemm.cc:383-391,
 * src/matrix.cc:403).  acc may be NULL. */
void orc_gemm_u8(const uint8_t *a, float sa, int32_t zpa, const uint8_t *b, float sb,
                 int32_t zpb, int m, int n, int k, float *c, int32_t *acc) {
  int i, j, kk;
  float c_scale = sa * sb;
  int32_t *row = (int32_t *)malloc(sizeof(int32_t) * n);
  for (i = 0; i < m; ++i) {
    for (j = 0; j < n; ++j) row[j] = 0;
    for (kk = 0; kk < k; ++kk) {
      int32_t av = (int32_t)a[(size_t)i * k + kk] - zpa;
      const uint8_t *brow = b + (size_t)kk * n;
      for (j = 0; j < n; ++j) row[j] += av * ((int32_t)brow[j] - zpb);
    }
    for (j = 0; j < n; ++j) {
      if (acc) acc[(size_t)i * n + j] = row[j];
      c[(size_t)i * n + j] = (float)row[j] * c_scale;
    }
  }
  free(row);
}

/* ------------------------------------------------------------------------- */
/* Nnet layers (src/nnet.cc) and the NN02 reader                              */
/* ------------------------------------------------------------------------- */

enum { L_LINEAR = 0, L_RELU = 1, L_NORMALIZE = 2, L_SOFTMAX = 3, L_SPLICE = 6,
       L_BATCHNORM = 7, L_LOGSOFTMAX = 8, L_NARROW = 9 };   /* src/nnet.h:21-30 */

typedef struct {
  int type;
  int in_dim, out_dim;      /* Linear */
  float *W, *b;             /* W [in x out] */
  uint8_t *W8; float w_scale; int32_t w_zp;
  int n_idx; int *idx;      /* Splice */
  int left, right;          /* Narrow */
  float *scale, *offset; int bn_dim;   /* BatchNorm */
} orc_layer_t;

typedef struct {
  int n_layers, left, right;
  orc_layer_t *layers;
} orc_nnet_t;

static int read_i32(FILE *f, int32_t *v) { return fread(v, 4, 1, f) == 1 ? 0 : -1; }
static int read_tag(FILE *f, const char *tag) {
  char b[4];
  if (fread(b, 1, 4, f) != 4) return -1;
  return memcmp(b, tag, 4) == 0 ? 0 : -1;
}
/* VEC0 (src/vector.cc:267-300) */
static float *read_vec(FILE *f, int *dim) {
  int32_t bytes, d;
  float *v;
  if (read_tag(f, "VEC0") || read_i32(f, &bytes) || read_i32(f, &d)) return NULL;
  if (d * 4 + 4 != bytes) return NULL;
  v = (float *)malloc(sizeof(float) * (d > 0 ? d : 1));
  if (fread(v, 4, d, f) != (size_t)d) { free(v); return NULL; }
  *dim = d;
  return v;
}
/* MAT0 (src/matrix.cc:160-191) */
static float *read_mat(FILE *f, int *rows, int *cols) {
  int32_t sz, r, c, i;
  float *m;
  if (read_tag(f, "MAT0") || read_i32(f, &sz) || read_i32(f, &r) || read_i32(f, &c)) return NULL;
  m = (float *)malloc(sizeof(float) * (size_t)r * c);
  for (i = 0; i < r; ++i) {
    int d;
    float *row = read_vec(f, &d);
    if (!row || d != c) { free(row); free(m); return NULL; }
    memcpy(m + (size_t)i * c, row, sizeof(float) * c);
    free(row);
  }
  *rows = r; *cols = c;
  return m;
}

void orc_nnet_close(orc_nnet_t *nn) {
  int i;
  if (!nn) return;
  for (i = 0; i < nn->n_layers; ++i) {
    orc_layer_t *L = &nn->layers[i];
    free(L->W); free(L->b); free(L->W8); free(L->idx); free(L->scale); free(L->offset);
  }
  free(nn->layers);
  free(nn);
}

/* NN02 / LAY0 (src/nnet.cc:221-293). Weights are also quantised once. */
orc_nnet_t *orc_nnet_open(const char *path) {
  FILE *f = fopen(path, "rb");
  orc_nnet_t *nn;
  int32_t l, r, n, i;
  if (!f) return NULL;
  if (read_tag(f, "NN02") || read_i32(f, &l) || read_i32(f, &r) || read_i32(f, &n)) {
    fclose(f);
    return NULL;
  }
  nn = (orc_nnet_t *)calloc(1, sizeof(orc_nnet_t));
  nn->left = l; nn->right = r; nn->n_layers = n;
  nn->layers = (orc_layer_t *)calloc(n > 0 ? n : 1, sizeof(orc_layer_t));
  for (i = 0; i < n; ++i) {
    orc_layer_t *L = &nn->layers[i];
    int32_t type;
    int ok = 1;
    if (read_tag(f, "LAY0") || read_i32(f, &type)) { ok = 0; }
    L->type = type;
    if (ok) switch (type) {
      case L_LINEAR: {
        int bd;
        L->W = read_mat(f, &L->in_dim, &L->out_dim);
        L->b = L->W ? read_vec(f, &bd) : NULL;
        if (!L->W || !L->b || bd != L->out_dim) { ok = 0; break; }
        L->W8 = (uint8_t *)malloc((size_t)L->in_dim * L->out_dim);
        orc_quantize(L->W, (long)L->in_dim * L->out_dim, L->W8, &L->w_scale, &L->w_zp);
        break;
      }
      case L_SPLICE: {
        int32_t k, j;
        if (read_i32(f, &k) || k < 0) { ok = 0; break; }
        L->n_idx = k;
        L->idx = (int *)malloc(sizeof(int) * (k > 0 ? k : 1));
        for (j = 0; j < k; ++j) { int32_t v; if (read_i32(f, &v)) ok = 0; L->idx[j] = v; }
        break;
      }
      case L_NARROW: {
        int32_t a, b;
        if (read_i32(f, &a) || read_i32(f, &b)) { ok = 0; break; }
        L->left = a; L->right = b;
        break;
      }
      case L_BATCHNORM: {
        int d2;
        L->scale = read_vec(f, &L->bn_dim);
        L->offset = L->scale ? read_vec(f, &d2) : NULL;
        if (!L->scale || !L->offset || d2 != L->bn_dim) ok = 0;
        break;
      }
      case L_RELU: case L_NORMALIZE: case L_SOFTMAX: case L_LOGSOFTMAX: break;
      default: ok = 0;
    }
    if (!ok) { nn->n_layers = i + 1; fclose(f); orc_nnet_close(nn); return NULL; }
  }
  fclose(f);
  return nn;
}

int orc_nnet_contexts(const orc_nnet_t *nn, int *left, int *right) {
  *left = nn->left; *right = nn->right;
  return nn->n_layers;
}

/* Nnet::Propagate (src/nnet.cc:295-307).  mode 0: float Linear layers
 * (LinearLayer::Propagate :22-36); mode 1: the int8 composition of SURVEY D3
 * (Quantize(in) -> MatMat_U8U8F32 -> + b).  Returns a malloc()ed [rows x cols]
 * matrix.  dump_linear >= 0 (mode 1): copies that Linear layer's int32
 * accumulators into acc_out (caller-sized). */
float *orc_nnet_propagate(const orc_nnet_t *nn, const float *in, int rows, int cols, int mode,
                          int *out_rows, int *out_cols, int dump_linear, int32_t *acc_out) {
  float *cur = (float *)malloc(sizeof(float) * (size_t)rows * cols);
  int li, r, c, linear_ordinal = 0;
  memcpy(cur, in, sizeof(float) * (size_t)rows * cols);
  for (li = 0; li < nn->n_layers; ++li) {
    const orc_layer_t *L = &nn->layers[li];
    float *next = NULL;
    int nrows = rows, ncols = cols;
    switch (L->type) {
      case L_LINEAR: {
        ncols = L->out_dim;
        next = (float *)malloc(sizeof(float) * (size_t)rows * ncols);
        if (mode == 0) {
          orc_sgemm(cur, L->W, rows, ncols, cols, next);
        } else {
          uint8_t *q = (uint8_t *)malloc((size_t)rows * cols);
          float sa; int32_t za;
          int32_t *acc = (linear_ordinal == dump_linear) ? acc_out : NULL;
          orc_quantize(cur, (long)rows * cols, q, &sa, &za);
          orc_gemm_u8(q, sa, za, L->W8, L->w_scale, L->w_zp, rows, ncols, cols, next, acc);
          free(q);
        }
        for (r = 0; r < rows; ++r)
          for (c = 0; c < ncols; ++c) next[(size_t)r * ncols + c] += L->b[c];
        ++linear_ordinal;
        break;
      }
      case L_SPLICE: {                                   /* nnet.cc:50-75 */
        int t;
        ncols = cols * L->n_idx;
        next = (float *)malloc(sizeof(float) * (size_t)rows * ncols);
        for (r = 0; r < rows; ++r)
          for (t = 0; t < L->n_idx; ++t) {
            int src = r + L->idx[t];
            if (src < 0) src = 0;
            if (src > rows - 1) src = rows - 1;
            memcpy(next + (size_t)r * ncols + (size_t)t * cols, cur + (size_t)src * cols,
                   sizeof(float) * cols);
          }
        break;
      }
      case L_NARROW: {                                   /* nnet.cc:182-202 */
        if (rows <= L->left + L->right) {                /* passthrough branch :186-189 */
          next = (float *)malloc(sizeof(float) * (size_t)rows * cols);
          memcpy(next, cur, sizeof(float) * (size_t)rows * cols);
        } else {
          nrows = rows - L->left - L->right;
          next = (float *)malloc(sizeof(float) * (size_t)nrows * cols);
          memcpy(next, cur + (size_t)L->left * cols, sizeof(float) * (size_t)nrows * cols);
        }
        break;
      }
      case L_RELU:                                       /* nnet.cc:149-160 */
        next = cur; cur = NULL;
        for (r = 0; r < rows * cols; ++r) if (next[r] < 0.0f) next[r] = 0.0f;
        break;
      case L_BATCHNORM:                                  /* nnet.cc:106-117 */
        next = cur; cur = NULL;
        for (r = 0; r < rows; ++r)
          for (c = 0; c < cols; ++c) {
            float v = next[(size_t)r * cols + c];
            v *= L->scale[c];
            v += L->offset[c];
            next[(size_t)r * cols + c] = v;
          }
        break;
      case L_NORMALIZE:                                  /* nnet.cc:162-175 */
        next = cur; cur = NULL;
        for (r = 0; r < rows; ++r) {
          float *row = next + (size_t)r * cols, ss = 0.0f, D = (float)cols, sc;
          for (c = 0; c < cols; ++c) ss += row[c] * row[c];
          sc = (float)sqrt(D / (double)ss);
          for (c = 0; c < cols; ++c) row[c] *= sc;
        }
        break;
      case L_SOFTMAX:                                    /* nnet.cc:126-135, vector.cc:95-107 */
        next = cur; cur = NULL;
        for (r = 0; r < rows; ++r) {
          float *row = next + (size_t)r * cols, sum = 0.0f;
          for (c = 0; c < cols; ++c) { row[c] = expf(row[c]); sum += row[c]; }
          for (c = 0; c < cols; ++c) row[c] /= sum;
        }
        break;
      case L_LOGSOFTMAX:                                 /* nnet.cc:137-146, vector.cc:110-122 */
        next = cur; cur = NULL;
        for (r = 0; r < rows; ++r) {
          float *row = next + (size_t)r * cols, sum = 0.0f, lse;
          for (c = 0; c < cols; ++c) sum += expf(row[c]);   /* no max subtraction */
          lse = logf(sum);
          for (c = 0; c < cols; ++c) row[c] -= lse;
        }
        break;
      default:
        free(cur);
        return NULL;
    }
    free(cur);
    cur = next; rows = nrows; cols = ncols;
  }
  *out_rows = rows; *out_cols = cols;
  return cur;
}

void orc_free(void *p) { free(p); }

/* AcousticModel for one whole utterance evaluated as ONE batch (chunk_size > T,
 * SURVEY Q12): replicate-pad `left` copies of frame 0 and `right` copies of the
 * last frame (src/am.cc:119-124,152-155), Nnet::Propagate, then subtract
 * log(prior) (src/am.cc:43-44,109-112).  `prior` holds probabilities.
 * Returns malloc()ed [T x num_pdfs] or NULL. */
float *orc_am_forward(const orc_nnet_t *nn, const float *prior, int num_pdfs, int left,
                      int right, const float *feats, int T, int dim, int mode, int *out_rows,
                      int *out_cols, int dump_linear, int32_t *acc_out) {
  int rows = T + left + right, r, c;
  float *in, *out;
  if (T <= 0) return NULL;
  in = (float *)malloc(sizeof(float) * (size_t)rows * dim);
  for (r = 0; r < rows; ++r) {
    int src = r - left;
    if (src < 0) src = 0;
    if (src > T - 1) src = T - 1;
    memcpy(in + (size_t)r * dim, feats + (size_t)src * dim, sizeof(float) * dim);
  }
  out = orc_nnet_propagate(nn, in, rows, dim, mode, out_rows, out_cols, dump_linear, acc_out);
  free(in);
  if (!out) return NULL;
  if (*out_cols != num_pdfs) { free(out); return NULL; }
  for (r = 0; r < *out_rows; ++r)
    for (c = 0; c < num_pdfs; ++c) {
      float lp = logf(prior[c]);
      out[(size_t)r * num_pdfs + c] += -1.0f * lp;
    }
  return out;
}
