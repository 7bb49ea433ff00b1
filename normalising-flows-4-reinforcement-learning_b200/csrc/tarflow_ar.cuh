// Persistent autoregressive-inverse kernel of the causal-transformer flow for evaluation-sized batches: see tarflow_ar.cu.
#pragma once
#include "common.cuh"

namespace nf {

struct ArShape { int B, A, C, EC, L, hd; };

struct ArP {
  // parameters of ONE block (device pointers into the flat arena)
  const float *pos, *win, *bin, *wout, *bout;
  const float* layers; long long layer_stride;      // layer 0 and the stride between layers
  long long o_ln1w, o_ln1b, o_wqkv, o_bqkv, o_wproj, o_bproj, o_ln2w, o_ln2b, o_w1, o_b1, o_w2, o_b2;   // offsets inside a layer
  int A, C, H, hd, L, EC, B, flip, ld_first;
  float eps, clamp;
  const float* zin;      // (B, A) this block's z
  float* xout;           // (B, A) this block's x
  float* logdet;         // (B) accumulated, may be null
  const float* cproj;    // (B, C) cond_proj(y) of this block
  float* qkv;            // (L, B, A, 3C) K / V cache (and the step's q)
  float *o, *h, *h2, *g; // (B, C) x3 and (B, EC): activations exchanged between the phases (L2-resident)
  unsigned* bar;         // grid-barrier counter, zero on entry
  long long* dbg;        // optional: 8 cycle counters per CTA (embed, P1..P5, tail, barriers); nf_debug_tc_timeline buffer
  int cp, w_smem;        // filled by tf_ar_block: columns per CTA, weights resident in shared memory
};

// true when the kernel can run this shape on the current device (B <= 128, widths multiples of 32 up to 1024, cooperative
// launch available, shared memory fits)
bool tf_ar_plan(const ArShape& s, int* cp_out, int* w_smem_out);
int tf_ar_block(ArP p, cudaStream_t st);

}  // namespace nf
