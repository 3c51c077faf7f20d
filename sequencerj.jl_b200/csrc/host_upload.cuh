// Host -> device upload of the caller's matrix.  The reference's caller holds a plain Julia `Matrix{Float64}`
// (src/sequencer.jl:105), i.e. PAGEABLE memory: one cudaMemcpyAsync from it runs at ~12 GB/s through the driver's
// single-threaded staging copy (28 ms for the 328 MB of BASELINE config 3, against 5.9 ms from pinned memory).
// Pageable inputs therefore go through the library's own pinned staging: a few host threads copy slabs into pinned
// buffers (the host-side memcpy is the bottleneck, so it is the part that is parallelised) and each slab is sent with
// an asynchronous copy as soon as it is staged, overlapping the next slab's memcpy.  Pinned / registered inputs
// (cudaHostAlloc, cudaHostRegister, torch pin_memory) are copied directly.  Included once, by seqj_api.cu.
#pragma once

namespace {

constexpr size_t kUploadSlab = size_t(8) << 20;      // bytes per staged slab
constexpr int kUploadMaxThreads = 8;
constexpr int kUploadBufsPerThread = 2;

struct UploadStage {
  void* pinned = nullptr;                            // kUploadMaxThreads * kUploadBufsPerThread slabs
  cudaEvent_t ev[kUploadMaxThreads * kUploadBufsPerThread] = {};
  bool ready = false;
};

inline bool host_pointer_is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

inline cudaError_t upload_stage_init(UploadStage& u) {
  if (u.ready) return cudaSuccess;
  cudaError_t e = cudaHostAlloc(&u.pinned, kUploadSlab * kUploadMaxThreads * kUploadBufsPerThread, cudaHostAllocDefault);
  if (e != cudaSuccess) return e;
  for (auto& ev : u.ev) {
    e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  u.ready = true;
  return cudaSuccess;
}

inline void upload_stage_free(UploadStage& u) {
  if (u.pinned) cudaFreeHost(u.pinned);
  for (auto& ev : u.ev) if (ev) cudaEventDestroy(ev);
  u = UploadStage();
}

// Copies `bytes` from pageable `src` to device `dst` on `stream`; returns after every slab has been ENQUEUED (the
// staging buffers are only reused after their previous copy has completed).  `dev` = device of the stream.
inline cudaError_t upload_pageable(UploadStage& u, int dev, void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  cudaError_t e = upload_stage_init(u);
  if (e != cudaSuccess) return e;
  const size_t nslab = (bytes + kUploadSlab - 1) / kUploadSlab;
  int nth = (int)std::min<size_t>(kUploadMaxThreads, std::max<size_t>(1, nslab));
  const unsigned hc = std::thread::hardware_concurrency();
  if (hc) nth = std::max(1, std::min<int>(nth, (int)hc));
  if (const char* env = getenv("SEQJ_UPLOAD_THREADS")) nth = std::max(1, std::min(kUploadMaxThreads, atoi(env)));
  std::vector<cudaError_t> errs(nth, cudaSuccess);
  auto work = [&](int t) {
    cudaSetDevice(dev);
    int use = 0;
    for (size_t sidx = (size_t)t; sidx < nslab; sidx += (size_t)nth, ++use) {
      const int b = t * kUploadBufsPerThread + (use % kUploadBufsPerThread);
      char* stage = static_cast<char*>(u.pinned) + (size_t)b * kUploadSlab;
      if (use >= kUploadBufsPerThread) {
        cudaError_t we = cudaEventSynchronize(u.ev[b]);
        if (we != cudaSuccess) { errs[t] = we; return; }
      }
      const size_t off = sidx * kUploadSlab, n = std::min(kUploadSlab, bytes - off);
      std::memcpy(stage, static_cast<const char*>(src) + off, n);
      cudaError_t ce = cudaMemcpyAsync(static_cast<char*>(dst) + off, stage, n, cudaMemcpyHostToDevice, stream);
      if (ce == cudaSuccess) ce = cudaEventRecord(u.ev[b], stream);
      if (ce != cudaSuccess) { errs[t] = ce; return; }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nth; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  for (cudaError_t x : errs) if (x != cudaSuccess) return x;
  return cudaSuccess;
}

}  // namespace
