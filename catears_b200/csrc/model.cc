// model.cc -- readers for the reference's on-disk formats and the layer-program compiler.
//
// Formats (all little-endian, host-native):
//   VEC0 | int32 bytes(=4*dim+4) | int32 dim | data            src/vector.cc:267-300
//   MAT0 | int32(8) | int32 rows | int32 cols | rows x VEC0    src/matrix.cc:160-191
//   NN02 | int32 L | int32 R | int32 n | n x (LAY0 | int32 type | payload)
//                                                              src/nnet.cc:221-293
//   Splice payload: int32 n + n x int32 (src/nnet.cc:77-95); Narrow: 2 x int32 (:204-215);
//   BatchNorm: 2 x VEC0 (:119-124); Linear: MAT0 W [in x out] + VEC0 b.
#include "model.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

namespace ce {
namespace {

class File {
 public:
  ~File() { if (f_) fclose(f_); }
  int Open(const std::string &path) {
    path_ = path;
    f_ = fopen(path.c_str(), "rb");
    if (!f_) {
      SetError("unable to open %s", path.c_str());
      return CE_GPU_EIO;
    }
    if (fseek(f_, 0, SEEK_END) == 0) {
      size_ = ftell(f_);
      fseek(f_, 0, SEEK_SET);
    }
    return CE_GPU_OK;
  }
  // true if `bytes` more bytes can still be read (dimensions in a corrupt header must not drive an
  // allocation the file cannot back)
  bool Has(uint64_t bytes) const {
    if (size_ < 0) return true;
    const long pos = ftell(f_);
    return pos >= 0 && bytes <= (uint64_t)(size_ - pos);
  }
  int Read(void *dst, size_t n) {
    if (n && fread(dst, 1, n, f_) != n) {
      SetError("%s: unexpected end of file", path_.c_str());
      return CE_GPU_EIO;
    }
    return CE_GPU_OK;
  }
  int I32(int32_t *v) { return Read(v, 4); }
  int Tag(const char *tag) {
    char b[4];
    CE_CHECK(Read(b, 4));
    if (memcmp(b, tag, 4) != 0) {
      SetError("%s: section name mismatch: expected %s, found %.4s", path_.c_str(), tag, b);
      return CE_GPU_EIO;
    }
    return CE_GPU_OK;
  }
  const std::string &path() const { return path_; }

 private:
  FILE *f_ = nullptr;
  long size_ = -1;
  std::string path_;
};

int ReadVec(File *f, std::vector<float> *v) {
  CE_CHECK(f->Tag("VEC0"));
  int32_t bytes = 0, dim = 0;
  CE_CHECK(f->I32(&bytes));
  CE_CHECK(f->I32(&dim));
  if (dim < 0 || bytes != 4 * dim + 4 || !f->Has(4ull * (uint64_t)dim)) {
    SetError("%s: VEC0 section size mismatch (%d bytes for dim %d)", f->path().c_str(), bytes, dim);
    return CE_GPU_EIO;
  }
  v->resize(dim);
  return f->Read(v->data(), sizeof(float) * (size_t)dim);
}

int ReadMat(File *f, std::vector<float> *m, int *rows, int *cols) {
  CE_CHECK(f->Tag("MAT0"));
  int32_t sz = 0, r = 0, c = 0;
  CE_CHECK(f->I32(&sz));
  CE_CHECK(f->I32(&r));
  CE_CHECK(f->I32(&c));
  if (sz != 8 || r < 0 || c < 0 || !f->Has((uint64_t)r * (12ull + 4ull * (uint64_t)c))) {
    SetError("%s: MAT0 header corrupt (size %d, %d x %d)", f->path().c_str(), sz, r, c);
    return CE_GPU_EIO;
  }
  m->resize((size_t)r * c);
  std::vector<float> row;
  for (int i = 0; i < r; ++i) {
    CE_CHECK(ReadVec(f, &row));
    if ((int)row.size() != c) {
      SetError("%s: MAT0 row %d has %zu columns, expected %d", f->path().c_str(), i, row.size(), c);
      return CE_GPU_EIO;
    }
    memcpy(m->data() + (size_t)i * c, row.data(), sizeof(float) * c);
  }
  *rows = r;
  *cols = c;
  return CE_GPU_OK;
}

std::string Trim(const std::string &s) {
  size_t b = s.find_first_not_of(" \t\r\n");
  if (b == std::string::npos) return "";
  size_t e = s.find_last_not_of(" \t\r\n");
  return s.substr(b, e - b + 1);
}

}  // namespace

int ReadVectorFile(const std::string &path, std::vector<float> *v) {
  File f;
  CE_CHECK(f.Open(path));
  return ReadVec(&f, v);
}

int ReadNnetFile(const std::string &path, HostNnet *nn) {
  File f;
  CE_CHECK(f.Open(path));
  CE_CHECK(f.Tag("NN02"));
  int32_t n = 0;
  CE_CHECK(f.I32(&nn->left_context));
  CE_CHECK(f.I32(&nn->right_context));
  CE_CHECK(f.I32(&n));
  if (n < 0 || n > 100000) {
    SetError("%s: implausible layer count %d", path.c_str(), n);
    return CE_GPU_EIO;
  }
  nn->layers.clear();
  nn->layers.resize(n);
  for (int i = 0; i < n; ++i) {
    HostLayer &L = nn->layers[i];
    CE_CHECK(f.Tag("LAY0"));
    int32_t type = 0;
    CE_CHECK(f.I32(&type));
    L.type = type;
    switch (type) {
      case kLinear: {
        CE_CHECK(ReadMat(&f, &L.W, &L.in_dim, &L.out_dim));
        CE_CHECK(ReadVec(&f, &L.b));
        if ((int)L.b.size() != L.out_dim) {
          SetError("%s: layer %d: bias has %zu entries for %d outputs", path.c_str(), i, L.b.size(),
                   L.out_dim);
          return CE_GPU_EIO;
        }
        break;
      }
      case kSplice: {
        int32_t k = 0;
        CE_CHECK(f.I32(&k));
        if (k < 1 || k > 1024) {
          SetError("%s: layer %d: splice with %d indices", path.c_str(), i, k);
          return CE_GPU_EIO;
        }
        L.indices.resize(k);
        CE_CHECK(f.Read(L.indices.data(), 4 * (size_t)k));
        break;
      }
      case kNarrow: {
        int32_t l = 0, r = 0;
        CE_CHECK(f.I32(&l));
        CE_CHECK(f.I32(&r));
        L.left = l;
        L.right = r;
        break;
      }
      case kBatchNorm: {
        CE_CHECK(ReadVec(&f, &L.scale));
        CE_CHECK(ReadVec(&f, &L.offset));
        if (L.scale.size() != L.offset.size()) {
          SetError("%s: layer %d: batch-norm scale/offset sizes differ", path.c_str(), i);
          return CE_GPU_EIO;
        }
        break;
      }
      case kReLU:
      case kNormalize:
      case kSoftmax:
      case kLogSoftmax:
        break;
      default:
        SetError("%s: unexpected layer type: %d", path.c_str(), type);   // nnet.cc:266
        return CE_GPU_EIO;
    }
  }
  return CE_GPU_OK;
}

int ReadConfigFile(const std::string &path, std::map<std::string, std::string> *kv,
                   std::string *dir) {
  FILE *f = fopen(path.c_str(), "r");
  if (!f) {
    SetError("unable to open %s", path.c_str());
    return CE_GPU_EIO;
  }
  char buf[4096];
  int rc = CE_GPU_OK;
  while (fgets(buf, sizeof(buf), f)) {
    if (strlen(buf) == sizeof(buf) - 1 && buf[sizeof(buf) - 2] != '\n') {
      SetError("%s: line longer than %zu bytes", path.c_str(), sizeof(buf) - 2);
      rc = CE_GPU_EIO;
      break;
    }
    std::string line = Trim(buf);
    if (line.empty() || line[0] == '#') continue;
    size_t eq = line.find('=');
    if (eq == std::string::npos || line.find('=', eq + 1) != std::string::npos) {
      SetError("Unexpected line in %s: %s", path.c_str(), line.c_str());   // configuration.cc:35
      rc = CE_GPU_EIO;
      break;
    }
    std::string key = Trim(line.substr(0, eq)), val = Trim(line.substr(eq + 1));
    std::transform(key.begin(), key.end(), key.begin(), ::tolower);
    if (val.empty()) {
      SetError("Value could not be empty: %s: %s", path.c_str(), line.c_str());
      rc = CE_GPU_EIO;
      break;
    }
    (*kv)[key] = val;
  }
  fclose(f);
  size_t pos = path.rfind('/');
  *dir = pos == std::string::npos ? "" : path.substr(0, pos + 1);
  return rc;
}

namespace {

// Any layer list: one step per layer.  The reference's matrices shrink at every NarrowLayer; here
// the rows stay where they are and the valid range [lo, P - hi) of every utterance block shrinks.
int CompileGeneral(const HostNnet &nn, int left_context, int right_context, int num_out, Program *prog) {
  prog->blocks.clear();
  prog->steps.clear();
  prog->general = true;
  prog->log_softmax = false;
  const int n = (int)nn.layers.size();
  if (n == 0) {
    SetError("the nnet has no layers");
    return CE_GPU_EUNSUPPORTED;
  }
  // feature dimension: the first layer that pins a width, divided by the splices in front of it
  int64_t mult = 1;
  int feat = 0;
  for (int i = 0; i < n && feat == 0; ++i) {
    const HostLayer &L = nn.layers[i];
    if (L.type == kSplice) mult *= (int64_t)L.indices.size();
    const int pinned = L.type == kLinear ? L.in_dim : L.type == kBatchNorm ? (int)L.scale.size() : 0;
    if (pinned > 0) {
      if (pinned % mult != 0) {
        SetError("layer %d expects %d inputs, not a multiple of the %lld spliced copies in front of it", i,
                 pinned, (long long)mult);
        return CE_GPU_EINVAL;
      }
      feat = (int)(pinned / mult);
    }
  }
  if (feat == 0) {
    if (num_out <= 0 || num_out % mult != 0) {
      SetError("the nnet has no Linear or BatchNorm layer and the prior (%d entries) does not pin its "
               "input dimension", num_out);
      return CE_GPU_EINVAL;
    }
    feat = (int)(num_out / mult);
  }
  prog->feat_dim = feat;
  int dim = feat, lo = 0, hi = 0;
  prog->max_dim = dim;
  for (int i = 0; i < n; ++i) {
    const HostLayer &L = nn.layers[i];
    if (L.type == kLogSoftmax && i == n - 1) {           // fused with the prior subtraction + argmax
      prog->log_softmax = true;
      break;
    }
    Step st;
    st.type = L.type;
    st.layer = i;
    st.in_dim = dim;
    st.lo = lo;
    st.hi = hi;
    switch (L.type) {
      case kLinear: {
        if (L.in_dim != dim) {
          SetError("layer %d: Linear expects %d inputs, previous layer produces %d", i, L.in_dim, dim);
          return CE_GPU_EINVAL;
        }
        Block b;
        b.taps.assign(1, 0);
        b.in_dim = L.in_dim;
        b.out_dim = L.out_dim;
        b.linear = i;
        b.cum_left = lo;
        b.cum_right = hi;
        st.block = (int)prog->blocks.size();
        prog->blocks.push_back(b);
        dim = L.out_dim;
        break;
      }
      case kSplice:
        if ((int64_t)dim * (int64_t)L.indices.size() > (1 << 20)) {
          SetError("layer %d: a spliced row of %lld floats is not supported", i,
                   (long long)dim * (long long)L.indices.size());
          return CE_GPU_EUNSUPPORTED;
        }
        dim *= (int)L.indices.size();
        break;
      case kNarrow:
        if (L.left < 0 || L.right < 0) {
          SetError("layer %d: NarrowLayer is not initialized", i);   // nnet.cc:185
          return CE_GPU_EINVAL;
        }
        lo += L.left;
        hi += L.right;
        break;
      case kBatchNorm:
        if ((int)L.scale.size() != dim) {
          SetError("layer %d: batch-norm dim %zu != %d", i, L.scale.size(), dim);
          return CE_GPU_EINVAL;
        }
        break;
      case kReLU: case kNormalize: case kSoftmax: case kLogSoftmax:
        break;
      default:
        SetError("layer %d: unexpected layer type %d", i, L.type);
        return CE_GPU_EUNSUPPORTED;
    }
    st.out_dim = dim;
    prog->max_dim = std::max(prog->max_dim, dim);
    prog->steps.push_back(st);
  }
  if (lo != left_context || hi != right_context) {
    // The reference would abort on assert(rows == batch_size), src/am.cc:106.
    SetError("left/right context %d/%d does not match the rows the nnet removes (%d/%d)",
             left_context, right_context, lo, hi);
    return CE_GPU_EINVAL;
  }
  prog->num_pdfs = dim;
  return CE_GPU_OK;
}

int CompileFused(const HostNnet &nn, int left_context, int right_context, Program *prog);

}  // namespace

int CompileProgram(const HostNnet &nn, int left_context, int right_context, Program *prog, int num_out) {
  const int rc = CompileFused(nn, left_context, right_context, prog);
  if (rc != CE_GPU_EUNSUPPORTED) return rc;
  ClearError();
  return CompileGeneral(nn, left_context, right_context, num_out, prog);
}

namespace {

int CompileFused(const HostNnet &nn, int left_context, int right_context, Program *prog) {
  prog->general = false;
  prog->steps.clear();
  prog->blocks.clear();
  prog->log_softmax = false;
  const int n = (int)nn.layers.size();
  int i = 0, cum_l = 0, cum_r = 0, dim = -1;
  while (i < n) {
    const HostLayer &L = nn.layers[i];
    if (L.type == kLogSoftmax && i == n - 1 && !prog->blocks.empty()) {
      prog->log_softmax = true;
      ++i;
      continue;
    }
    Block b;
    if (L.type == kSplice) {
      if (i + 2 >= n || nn.layers[i + 1].type != kNarrow || nn.layers[i + 2].type != kLinear) {
        SetError("layer %d: a Splice must be followed by Narrow and Linear (tool/convert_am.py "
                 "pattern); other stacks are not supported by the GPU program", i);
        return CE_GPU_EUNSUPPORTED;
      }
      b.taps.assign(L.indices.begin(), L.indices.end());
      if ((int)b.taps.size() > kMaxTaps) {
        SetError("layer %d: %zu splice indices (max %d)", i, b.taps.size(), kMaxTaps);
        return CE_GPU_EUNSUPPORTED;
      }
      int mn = 0, mx = 0;
      for (int32_t t : b.taps) {
        mn = std::min(mn, (int)t);
        mx = std::max(mx, (int)t);
      }
      const HostLayer &N = nn.layers[i + 1];
      if (N.left != -mn || N.right != mx) {
        SetError("layer %d: Narrow(%d,%d) does not remove exactly the clamped rows of "
                 "Splice[%d..%d]", i + 1, N.left, N.right, mn, mx);
        return CE_GPU_EUNSUPPORTED;
      }
      b.narrow_left = N.left;
      b.narrow_right = N.right;
      i += 2;
    } else if (L.type == kLinear) {
      b.taps.assign(1, 0);
    } else {
      SetError("layer %d (type %d) is not part of a [Splice,Narrow,]Linear[,ReLU][,BatchNorm] "
               "block; the GPU program does not support it", i, L.type);
      return CE_GPU_EUNSUPPORTED;
    }
    const HostLayer &Lin = nn.layers[i];
    b.linear = i;
    b.out_dim = Lin.out_dim;
    if (Lin.in_dim % (int)b.taps.size() != 0) {
      SetError("layer %d: Linear input %d is not a multiple of %zu splice taps", i, Lin.in_dim,
               b.taps.size());
      return CE_GPU_EINVAL;
    }
    b.in_dim = Lin.in_dim / (int)b.taps.size();
    if (dim >= 0 && b.in_dim != dim) {
      SetError("layer %d: Linear expects %d inputs per tap, previous layer produces %d", i, b.in_dim,
               dim);
      return CE_GPU_EINVAL;
    }
    if (dim < 0) prog->feat_dim = b.in_dim;
    ++i;
    if (i < n && nn.layers[i].type == kReLU) {
      b.relu = true;
      ++i;
    }
    if (i < n && nn.layers[i].type == kBatchNorm) {
      if ((int)nn.layers[i].scale.size() != b.out_dim) {
        SetError("layer %d: batch-norm dim %zu != %d", i, nn.layers[i].scale.size(), b.out_dim);
        return CE_GPU_EINVAL;
      }
      b.batchnorm = i;
      ++i;
    }
    cum_l += b.narrow_left;
    cum_r += b.narrow_right;
    b.cum_left = cum_l;
    b.cum_right = cum_r;
    dim = b.out_dim;
    prog->blocks.push_back(b);
  }
  if (prog->blocks.empty()) {
    SetError("the nnet has no Linear layer");
    return CE_GPU_EUNSUPPORTED;
  }
  if (cum_l != left_context || cum_r != right_context) {
    // The reference would abort on assert(rows == batch_size), src/am.cc:106.
    SetError("left/right context %d/%d does not match the rows the nnet removes (%d/%d)",
             left_context, right_context, cum_l, cum_r);
    return CE_GPU_EINVAL;
  }
  prog->num_pdfs = dim;
  return CE_GPU_OK;
}

}  // namespace

void QuantizeHost(const float *src, int64_t count, uint8_t *dst, float *scale_out,
                  int32_t *zp_out) {
  float mn = FLT_MAX, mx = FLT_MIN;                       // matrix.cc:330-331 (FLT_MIN quirk)
  for (int64_t i = 0; i < count; ++i) {
    float v = src[i];
    if (v > mx) mx = v;
    if (v < mn) mn = v;
  }
  volatile float range = mx - mn;                         // float subtraction, then double
  double scale = (double)range / 255.0;                   // matrix.cc:354
  double fzp = (double)(-mn) / scale;
  int32_t zp = (int32_t)round(fzp);
  float scale_f = (float)scale;
  for (int64_t i = 0; i < count; ++i) {
    volatile float q = src[i] / scale_f;                  // matrix.cc:383 (no contraction)
    float v = q + (float)zp;
    v = std::max(0.0f, std::min(v, 255.0f));
    dst[i] = (uint8_t)roundf(v);
  }
  *scale_out = scale_f;
  *zp_out = zp;
}

}  // namespace ce
