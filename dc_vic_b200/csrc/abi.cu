// extern "C" entry points that dispatch between the VQ kernels (see include/dcvic_b200.h).
#include "vq_common.cuh"

using namespace dcvic;

extern "C" const char* dcvic_version(void) { return "dcvic_b200 0.1 (sm_100a)"; }

extern "C" const char* dcvic_error_string(int code) {
  switch (code) {
    case DCVIC_OK: return "ok";
    case DCVIC_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or misaligned buffer)";
    case DCVIC_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
    case DCVIC_ERR_WORKSPACE: return "workspace missing or too small";
    case DCVIC_ERR_CUDA: return "CUDA error while enqueuing";
    case DCVIC_ERR_DEVICE: return "device is not sm_100";
    default: return "unknown error";
  }
}

static inline bool vq_narrow_ok(int D) { return D == 4 || D == 8; }

extern "C" int dcvic_vq_path(int D, int K, int flags) {
  if (D <= 0 || K <= 0) return DCVIC_ERR_BAD_ARG;
  if (!(flags & (DCVIC_VQ_FORCE_EXACT | DCVIC_VQ_RAGGED_HW)) && vq_tensor_supported(D, K)) return 2;
  if (flags & DCVIC_VQ_FORCE_TENSOR) return DCVIC_ERR_UNSUPPORTED;
  if (vq_narrow_ok(D) && !(flags & DCVIC_VQ_FORCE_EXACT)) return 0;
  return D <= 1024 ? 1 : DCVIC_ERR_UNSUPPORTED;
}

extern "C" size_t dcvic_vq_workspace_bytes(int B, int D, int H, int W, int K) {
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || K <= 0) return 0;
  if ((long long)B * H * W > 0x7fffffffLL) return 0;
  return vq_workspace_layout(B, D, H * W, K).total;
}

extern "C" int dcvic_vq_forward(const float* z_nchw, const float* codebook, int B, int D, int H, int W, int K,
                                float beta, int legacy, float* zq_nchw, int64_t* idx, float* loss, float* onehot,
                                float* perplexity, int flags, void* workspace, size_t ws_bytes,
                                dcvic_stream_t stream) {
  DCVIC_CHECK_ARG(z_nchw && codebook && zq_nchw && idx && loss && workspace);
  DCVIC_CHECK_ARG(B > 0 && D > 0 && H > 0 && W > 0 && K > 0);
  if ((long long)B * H * W > 0x7fffffffLL || (long long)K * D > 0x7fffffffLL) return DCVIC_ERR_UNSUPPORTED;
  const int HW = H * W, N = B * HW;
  // the tensor search reads 4 tokens per 16-byte load: needs H*W % 4 == 0 and a 16-byte aligned z
  const bool ragged = (HW & 3) != 0 || (reinterpret_cast<uintptr_t>(z_nchw) & 15) != 0;
  const int path = dcvic_vq_path(D, K, flags | (ragged ? DCVIC_VQ_RAGGED_HW : 0));
  if (path < 0) return path;
  const VqWorkspace w = vq_workspace_layout(B, D, HW, K);
  if (ws_bytes < w.total) return DCVIC_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 255) return DCVIC_ERR_BAD_ARG;
  char* ws = reinterpret_cast<char*>(workspace);
  unsigned* counters = reinterpret_cast<unsigned*>(ws + w.off_counters);
  float* ee = reinterpret_cast<float*>(ws + w.off_ee);
  float* nhee = reinterpret_cast<float*>(ws + w.off_nhee);
  float* emax = reinterpret_cast<float*>(ws + w.off_emax);
  double* partials = reinterpret_cast<double*>(ws + w.off_partials);
  unsigned* hist = reinterpret_cast<unsigned*>(ws + w.off_hist);
  int* cand = reinterpret_cast<int*>(ws + w.off_cand);
  VqMeta* meta = reinterpret_cast<VqMeta*>(ws + w.off_meta);
  uint2* list = reinterpret_cast<uint2*>(ws + w.off_list);
  __half* cb16 = reinterpret_cast<__half*>(ws + w.off_cb16);
  float* eperm = reinterpret_cast<float*>(ws + w.off_eperm);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = DCVIC_OK;

  if (path == 0) {
    rc = vq_narrow_forward(z_nchw, codebook, B, D, HW, K, beta, legacy, zq_nchw, idx, loss, partials, counters, s);
  } else {
    const bool prepared = !(flags & (DCVIC_VQ_REUSE_PREP | DCVIC_VQ_STAGE_FINISH_ONLY));
    if (prepared) {
      rc = vq_prepare_codebook(codebook, K, D, ee, nhee, emax, path == 2 ? cb16 : nullptr,
                               path == 2 ? eperm : nullptr, counters, s);
      if (rc) return rc;
    }
    if (flags & DCVIC_VQ_STAGE_PREP_ONLY) return rc;
    const bool do_search = !(flags & DCVIC_VQ_STAGE_FINISH_ONLY);
    const bool do_finish = !(flags & DCVIC_VQ_STAGE_SEARCH_ONLY);
    if (path == 2 && do_search && do_finish && !(flags & DCVIC_VQ_TWO_KERNELS) &&
        vq_fused_supported(z_nchw, zq_nchw, codebook, D, HW, K)) {
      rc = vq_fused_forward(z_nchw, eperm, ee, emax, cb16, B, D, HW, K, prepared, beta, legacy, zq_nchw, idx, loss,
                            partials, counters, s);
    } else if (path == 2) {
      if (do_search)
        rc = vq_tensor_search(z_nchw, cb16, emax, B, D, HW, K, prepared,
                              do_finish && (reinterpret_cast<uintptr_t>(codebook) & 15) == 0 &&
                                  vq_finish_tma_supported(z_nchw, zq_nchw, D, HW, K),
                              meta, list, s);
      if (rc) return rc;
      if (do_finish)
        rc = vq_finish(z_nchw, codebook, ee, emax, nullptr, meta, list, B, D, HW, K, beta, legacy, zq_nchw, idx, loss,
                       partials, counters, s);
    } else {
      if (do_search) rc = vq_exact_search(z_nchw, codebook, ee, B, D, HW, K, cand, s);
      if (rc) return rc;
      if (do_finish)
        rc = vq_finish(z_nchw, codebook, ee, emax, cand, nullptr, nullptr, B, D, HW, K, beta, legacy, zq_nchw, idx,
                       loss, partials, counters, s);
    }
    if (!do_finish) return rc;
  }
  if (rc) return rc;
  if (onehot || perplexity) rc = vq_v1_extras(idx, N, K, onehot, perplexity, hist, counters, s);
  return rc;
}
