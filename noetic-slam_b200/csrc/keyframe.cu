// Device-resident keyframe store and submap assembly (SURVEY.md §8f "next" row 1; additive API).
// The reference keeps every keyframe's cloud and covariance list on the host, transforms them there
// (src/dlio/src/dlio/odom.cc:1757-1762: pcl::transformPointCloud + cov <- Td * cov * Td^T), concatenates the selected
// keyframes into the submap cloud and the submap covariance list (odom.cc:1719-1729) and hands both back through
// setInputTarget / setTargetCovariances (odom.cc:1737, :998) — a 128 B/point Matrix4d round trip (128 MB for a
// 1M-point submap). Here a keyframe is captured from the scan that was just registered (its points and covariances
// are already in HBM), transformed in place by a kernel, and a submap is a device-to-device concatenation followed by
// the normal index build. Nothing crosses PCIe.
#include <vector>

#include "internal.h"
#include "linearize.cuh"

struct ngicp_keyframe {
  int device = 0;
  size_t n = 0;
  float4* pts = nullptr;   // [n] original scan order, w unused
  float* cov6 = nullptr;   // [n][6] original scan order
};

namespace ngicp {
namespace {

__global__ void __launch_bounds__(256) capture_kernel(const float4* __restrict__ sorted_pts, const float* __restrict__ sorted_cov6, int n,
                                                      float4* __restrict__ pts, float* __restrict__ cov6) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float4 p = __ldg(sorted_pts + j);
  const int o = __float_as_int(p.w);
  pts[o] = make_float4(p.x, p.y, p.z, 1.0f);
  for (int i = 0; i < 6; i++) cov6[(size_t)o * 6 + i] = sorted_cov6[(size_t)j * 6 + i];
}

// points: fp32 R p + t (pcl::transformPointCloud, odom.cc:1757-1758); covariances: Td C Td^T with Td = T.cast<double>()
// (odom.cc:1760-1762), 3x3 block only (row/column 3 of the reference's Matrix4d stay zero).
__global__ void __launch_bounds__(256) transform_kernel(float4* __restrict__ pts, float* __restrict__ cov6, int n, PoseArg P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  float q[3];
#pragma unroll
  for (int r = 0; r < 3; r++)
    q[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.Rf[3 * r], p.x), __fmul_rn(P.Rf[3 * r + 1], p.y)), __fmul_rn(P.Rf[3 * r + 2], p.z)), P.tf[r]);
  pts[i] = make_float4(q[0], q[1], q[2], 1.0f);
  float* c = cov6 + (size_t)i * 6;
  const double a00 = c[0], a01 = c[1], a02 = c[2], a11 = c[3], a12 = c[4], a22 = c[5];
  double T[9];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const double r0 = P.R[3 * r], r1 = P.R[3 * r + 1], r2 = P.R[3 * r + 2];
    T[3 * r + 0] = r0 * a00 + r1 * a01 + r2 * a02;
    T[3 * r + 1] = r0 * a01 + r1 * a11 + r2 * a12;
    T[3 * r + 2] = r0 * a02 + r1 * a12 + r2 * a22;
  }
  const int ri[6] = {0, 0, 0, 1, 1, 2}, ci[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int e = 0; e < 6; e++) c[e] = (float)(T[3 * ri[e]] * P.R[3 * ci[e]] + T[3 * ri[e] + 1] * P.R[3 * ci[e] + 1] + T[3 * ri[e] + 2] * P.R[3 * ci[e] + 2]);
}

__global__ void __launch_bounds__(256) kf_to_host_layout_kernel(const float4* __restrict__ pts, const float* __restrict__ cov6, int n,
                                                                float* __restrict__ xyz, double* __restrict__ m16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (xyz) { const float4 p = pts[i]; xyz[3 * (size_t)i] = p.x; xyz[3 * (size_t)i + 1] = p.y; xyz[3 * (size_t)i + 2] = p.z; }
  if (m16) {
    const float* c = cov6 + (size_t)i * 6;
    double* o = m16 + (size_t)i * 16;
    o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = 0; o[4] = c[1]; o[5] = c[3]; o[6] = c[4]; o[7] = 0;
    o[8] = c[2]; o[9] = c[4]; o[10] = c[5]; o[11] = 0; o[12] = 0; o[13] = 0; o[14] = 0; o[15] = 0;
  }
}

}  // namespace
}  // namespace ngicp

using namespace ngicp;

extern "C" {

int ngicp_keyframe_capture(ngicp_handle* h, ngicp_keyframe** out) {
  if (!h || !out) return fail(h, NGICP_ERR_INVALID, "ngicp_keyframe_capture: NULL argument");
  *out = nullptr;
  if (int rc = select_device(h)) return rc;
  const Index* si = h->index[0];
  if (!si || si->n_seg != 1) return fail(h, NGICP_ERR_INVALID, "keyframe capture: no (single) source cloud");
  if (!h->covs[0].valid || h->covs[0].n != (size_t)si->n) return fail(h, NGICP_ERR_INVALID, "keyframe capture: source covariances missing");
  ngicp_keyframe* kf = new ngicp_keyframe;
  kf->device = h->device;
  kf->n = (size_t)si->n;
  NGICP_CUDA(h, dev_alloc(&kf->pts, kf->n, h->stream));
  NGICP_CUDA(h, dev_alloc(&kf->cov6, kf->n * 6, h->stream));
  capture_kernel<<<(si->n + 255) / 256, 256, 0, h->stream>>>(si->pts, h->covs[0].cov6, si->n, kf->pts, kf->cov6);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  *out = kf;
  return NGICP_OK;
}

int ngicp_keyframe_transform(ngicp_handle* h, ngicp_keyframe* kf, const float T[16]) {
  if (!h || !kf || !T) return fail(h, NGICP_ERR_INVALID, "ngicp_keyframe_transform: NULL argument");
  if (int rc = select_device(h)) return rc;
  double Td[16];
  for (int i = 0; i < 16; i++) Td[i] = (double)T[i];   // Td = T.cast<double>()
  PoseArg P = make_pose(Td);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) P.Rf[3 * r + c] = T[4 * c + r]; P.tf[r] = T[12 + r]; }
  transform_kernel<<<((int)kf->n + 255) / 256, 256, 0, h->stream>>>(kf->pts, kf->cov6, (int)kf->n, P);
  count_launch(h);
  NGICP_CUDA(h, cudaGetLastError());
  return NGICP_OK;
}

size_t ngicp_keyframe_size(const ngicp_keyframe* kf) { return kf ? kf->n : 0; }

int ngicp_keyframe_release(ngicp_handle* h, ngicp_keyframe* kf) {
  if (!kf) return NGICP_OK;
  cudaSetDevice(kf->device);
  cudaStream_t s = h ? h->stream : (cudaStream_t)0;
  if (h) cudaStreamSynchronize(h->stream);
  dev_free(kf->pts, s);
  dev_free(kf->cov6, s);
  delete kf;
  return NGICP_OK;
}

int ngicp_keyframe_download(ngicp_handle* h, const ngicp_keyframe* kf, float* xyz, double* cov_4x4) {
  if (!h || !kf) return fail(h, NGICP_ERR_INVALID, "ngicp_keyframe_download: NULL argument");
  if (int rc = select_device(h)) return rc;
  const int n = (int)kf->n;
  float* d_xyz = nullptr; double* d_m = nullptr;
  if (xyz) NGICP_CUDA(h, dev_alloc(&d_xyz, kf->n * 3, h->stream));
  if (cov_4x4) NGICP_CUDA(h, dev_alloc(&d_m, kf->n * 16, h->stream));
  kf_to_host_layout_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(kf->pts, kf->cov6, n, d_xyz, d_m);
  count_launch(h);
  if (xyz) NGICP_CUDA(h, cudaMemcpyAsync(xyz, d_xyz, sizeof(float) * 3 * kf->n, cudaMemcpyDeviceToHost, h->stream));
  if (cov_4x4) NGICP_CUDA(h, cudaMemcpyAsync(cov_4x4, d_m, sizeof(double) * 16 * kf->n, cudaMemcpyDeviceToHost, h->stream));
  NGICP_CUDA(h, cudaStreamSynchronize(h->stream));
  dev_free(d_xyz, h->stream); dev_free(d_m, h->stream);
  return NGICP_OK;
}

// buildSubmap (odom.cc:1719-1738): concatenate the chosen keyframes, in the given order, into the target cloud and
// its covariance list; build the index; everything stays on the device.
int ngicp_submap_assemble(ngicp_handle* h, ngicp_keyframe* const* kfs, int n_kfs) {
  if (!h || !kfs || n_kfs < 1) return fail(h, NGICP_ERR_INVALID, "ngicp_submap_assemble: bad argument");
  if (int rc = select_device(h)) return rc;
  size_t n = 0;
  for (int i = 0; i < n_kfs; i++) {
    if (!kfs[i] || kfs[i]->device != h->device) return fail(h, NGICP_ERR_INVALID, "ngicp_submap_assemble: bad keyframe");
    n += kfs[i]->n;
  }
  cudaStream_t s = h->stream;
  float4* d_pts = nullptr; float* d_cov = nullptr;
  NGICP_CUDA(h, dev_alloc(&d_pts, n, s));
  NGICP_CUDA(h, dev_alloc(&d_cov, n * 6, s));
  size_t off = 0;
  for (int i = 0; i < n_kfs; i++) {   // one plain async copy per keyframe and array
    NGICP_CUDA(h, cudaMemcpyAsync(d_pts + off, kfs[i]->pts, sizeof(float4) * kfs[i]->n, cudaMemcpyDeviceToDevice, s));
    NGICP_CUDA(h, cudaMemcpyAsync(d_cov + off * 6, kfs[i]->cov6, sizeof(float) * 6 * kfs[i]->n, cudaMemcpyDeviceToDevice, s));
    off += kfs[i]->n;
  }
  Index* idx = nullptr;
  int rc = build_index(h, reinterpret_cast<const float*>(d_pts), 4, (int)n, nullptr, 1, &idx, n <= 262144);   // submaps: no fine levels (covariances are the keyframes')
  dev_free(d_pts, s);
  if (rc) { dev_free(d_cov, s); return rc; }
  rc = swap_in_index(h, NGICP_TARGET, idx);   // drops the old target covariances, like setInputTarget
  if (rc) { dev_free(d_cov, s); return rc; }
  CovSet& c = h->covs[NGICP_TARGET];
  NGICP_CUDA(h, dev_alloc(&c.cov6, n * 6, s));
  rc = cov6_from_host_order(h, idx, d_cov, c.cov6);   // concatenation order -> the new index's sorted order
  dev_free(d_cov, s);
  if (rc) return rc;
  c.n = n; c.valid = true;
  return NGICP_OK;
}

}  // extern "C"
