// model.h -- host-side model: the reference's on-disk formats and the layer program the GPU runs.
#ifndef CE_GPU_MODEL_H_
#define CE_GPU_MODEL_H_

#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "gemm.h"

namespace ce {

// Layer type ids, src/nnet.h:21-30.
enum LayerType {
  kLinear = 0, kReLU = 1, kNormalize = 2, kSoftmax = 3, kSplice = 6, kBatchNorm = 7,
  kLogSoftmax = 8, kNarrow = 9
};

struct HostLayer {
  int type = -1;
  int in_dim = 0, out_dim = 0;        // Linear: W is [in x out] row-major (tool/convert_am.py:318-323)
  std::vector<float> W, b;
  std::vector<int32_t> indices;       // Splice
  int left = 0, right = 0;            // Narrow
  std::vector<float> scale, offset;   // BatchNorm
};

struct HostNnet {
  int left_context = 0, right_context = 0;   // header values (ignored by AM, src/am.cc:48-50)
  std::vector<HostLayer> layers;
};

// Readers for VEC0 / MAT0 / NN02 (src/vector.cc:267-300, src/matrix.cc:160-191,
// src/nnet.cc:221-293).  Return CE_GPU_EIO with a message on a missing / corrupt file.
int ReadVectorFile(const std::string &path, std::vector<float> *v);
int ReadNnetFile(const std::string &path, HostNnet *nn);
// key = value configuration file (src/configuration.cc:14-88): lower-cased keys, '#' comments.
int ReadConfigFile(const std::string &path, std::map<std::string, std::string> *kv,
                   std::string *dir);

// One fused GPU step: [Splice + Narrow] + Linear [+ ReLU] [+ BatchNorm].
struct Block {
  std::vector<int32_t> taps;          // row offsets (Splice indices); {0} without a Splice
  int narrow_left = 0, narrow_right = 0;
  int in_dim = 0;                     // channels per tap (C)
  int out_dim = 0;                    // N
  int linear = -1;                    // index into HostNnet::layers
  bool relu = false;
  int batchnorm = -1;                 // index into HostNnet::layers, -1 = none
  int cum_left = 0, cum_right = 0;    // rows of an utterance block that are valid AFTER this block:
                                      // [cum_left, P - cum_right)
};

// One step of the GENERAL program: layer `layer` of the HostNnet, evaluated on its own (Linear
// layers still run on the tensor cores, everything else as a row-wise kernel, layers.cu).
struct Step {
  int type = -1;                      // LayerType
  int layer = -1;                     // index into HostNnet::layers
  int block = -1;                     // Linear: index into Program::blocks
  int in_dim = 0, out_dim = 0;
  int lo = 0, hi = 0;                 // valid rows of an utterance block BEFORE this step: [lo, P - hi)
};

struct Program {
  std::vector<Block> blocks;
  bool log_softmax = false;           // trailing LogSoftmaxLayer
  int feat_dim = 0;
  int num_pdfs = 0;
  // general == false: `blocks` is the whole network as fused [Splice+Narrow+]Linear[+ReLU][+BN]
  // steps (the pattern tool/convert_am.py emits).  general == true: any layer list of
  // src/nnet.h:21-30 in `steps`, one kernel per layer; `blocks` then holds the bare Linear layers.
  bool general = false;
  std::vector<Step> steps;
  int max_dim = 0;                    // widest activation row of the general program
};

// Matches the layer list against the pattern tool/convert_am.py emits; any other
// stack (Normalize / Softmax layers, a Splice without its Narrow, ...) compiles to the general
// program.  num_out: the prior's length (pins the feature dimension of stacks without a Linear or
// BatchNorm layer; 0 = unknown).
int CompileProgram(const HostNnet &nn, int left_context, int right_context, Program *prog,
                   int num_out = 0);

// Quantize (src/matrix.cc:329-387) on the host, for weights at load time.
void QuantizeHost(const float *src, int64_t count, uint8_t *dst, float *scale, int32_t *zero_point);

}  // namespace ce

#endif  // CE_GPU_MODEL_H_
