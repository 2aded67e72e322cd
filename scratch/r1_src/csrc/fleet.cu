// aero-ddc-b200: the fleet object of include/aeroddc.h - one aeroddc_bank per GPU, VFOs sharded over them, the raw
// block broadcast with NCCL. Host orchestration only; all arithmetic is the banks'.
#include "../include/aeroddc.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

int aeroddc_set_error(int code, const char* fmt, ...);   // bank.cu

namespace {

// the few NCCL entry points used, resolved from libnccl.so.2 at first use (the library stays loadable without NCCL)
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
  int (*CommInitAll)(ncclComm_t*, int, const int*);
  int (*CommDestroy)(ncclComm_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  bool ok = false;
};
NcclApi& nccl() {
  static NcclApi a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      a.CommInitAll = (int (*)(ncclComm_t*, int, const int*))dlsym(h, "ncclCommInitAll");
      a.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
      a.Broadcast = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclBroadcast");
      a.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
      a.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
      a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
      a.ok = a.CommInitAll && a.CommDestroy && a.Broadcast && a.GroupStart && a.GroupEnd && a.GetErrorString;
    }
  }
  return a;
}
constexpr int kNcclUint8 = 1;   // ncclUint8 (nccl.h)

#define FCU(x)                                                                                             \
  do {                                                                                                     \
    cudaError_t e_ = (x);                                                                                  \
    if (e_ != cudaSuccess) return aeroddc_set_error(AERODDC_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define FNC(x)                                                                                             \
  do {                                                                                                     \
    int r_ = (x);                                                                                          \
    if (r_ != 0) return aeroddc_set_error(AERODDC_ERR_CUDA, "%s failed: %s", #x, nccl().GetErrorString(r_)); \
  } while (0)
#define FOK(x)              \
  do {                      \
    int r_ = (x);           \
    if (r_ < 0) return r_;  \
  } while (0)

struct Where { int dev, local; };

}  // namespace

struct aeroddc_fleet {
  int fs = 0, B = 0, fmt = 0;
  std::vector<int> devices;
  std::vector<aeroddc_bank*> banks;
  std::vector<Where> where;          // global VFO index -> (device slot, index in that bank)
  std::vector<int> per_dev;          // VFOs added per device (round-robin counter uses flat/main VFOs only)
  int next_flat = 0;
  bool finalized = false;
  size_t in_bytes = 0;
  std::vector<ncclComm_t> comms;
  std::vector<cudaStream_t> streams;            // per device: carries the broadcast
  std::vector<unsigned char*> d_in[2];          // per device raw block, two parities
  std::vector<cudaEvent_t> ev_ready[2];         // per device: block of parity p has arrived
  void* h_slot[2] = {nullptr, nullptr};
  long long submitted = 0, done = 0;
};

extern "C" {

int aeroddc_fleet_create(aeroddc_fleet** out, int sample_rate, int block_len, int in_format, const int* devices, int n_devices) {
  if (!out || !devices || n_devices < 1) return aeroddc_set_error(AERODDC_ERR_ARG, "bad fleet arguments");
  *out = nullptr;
  aeroddc_fleet* f = new aeroddc_fleet();
  f->fs = sample_rate; f->B = block_len; f->fmt = in_format;
  for (int i = 0; i < n_devices; ++i) {
    aeroddc_bank* b = nullptr;
    int rc = aeroddc_bank_create(&b, sample_rate, block_len, in_format, devices[i]);
    if (rc != AERODDC_OK) { aeroddc_fleet_destroy(f); return rc; }
    f->devices.push_back(devices[i]);
    f->banks.push_back(b);
    f->per_dev.push_back(0);
  }
  *out = f;
  return AERODDC_OK;
}

int aeroddc_fleet_add_vfo(aeroddc_fleet* f, const aeroddc_vfo_desc* d) {
  if (!f || !d) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL argument");
  if (f->finalized) return aeroddc_set_error(AERODDC_ERR_STATE, "fleet already finalized");
  aeroddc_vfo_desc local = *d;
  int dev;
  if (d->parent >= 0) {   // a sub-VFO lives where its main VFO lives (it reads that GPU's stage-D stream)
    if (d->parent >= (int)f->where.size()) return aeroddc_set_error(AERODDC_ERR_ARG, "parent %d is not an earlier VFO", d->parent);
    dev = f->where[d->parent].dev;
    local.parent = f->where[d->parent].local;
  } else {
    dev = f->next_flat++ % (int)f->banks.size();
  }
  const int li = aeroddc_bank_add_vfo(f->banks[dev], &local);
  if (li < 0) return li;
  f->where.push_back({dev, li});
  f->per_dev[dev]++;
  return (int)f->where.size() - 1;
}

int aeroddc_fleet_set_mode(aeroddc_fleet* f, int mode) {
  if (!f) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL fleet");
  for (aeroddc_bank* b : f->banks) FOK(aeroddc_bank_set_mode(b, mode));
  return AERODDC_OK;
}

int aeroddc_fleet_set_dc_correction(aeroddc_fleet* f, int enable) {
  if (!f) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL fleet");
  for (aeroddc_bank* b : f->banks) FOK(aeroddc_bank_set_dc_correction(b, enable));   // every GPU removes DC from its copy
  return AERODDC_OK;
}

int aeroddc_fleet_finalize(aeroddc_fleet* f) {
  if (!f) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL fleet");
  if (f->finalized) return aeroddc_set_error(AERODDC_ERR_STATE, "already finalized");
  const int n = (int)f->banks.size();
  // devices that received no VFO are dropped from the fleet (fewer VFOs than GPUs)
  for (int i = n - 1; i >= 1; --i)
    if (f->per_dev[i] == 0) {
      for (const Where& w : f->where) if (w.dev > i) return aeroddc_set_error(AERODDC_ERR_STATE, "internal: sparse device use");
      aeroddc_bank_destroy(f->banks[i]);
      f->banks.erase(f->banks.begin() + i); f->devices.erase(f->devices.begin() + i); f->per_dev.erase(f->per_dev.begin() + i);
    }
  if (f->per_dev[0] == 0) return aeroddc_set_error(AERODDC_ERR_STATE, "no VFOs");
  const int nd = (int)f->banks.size();
  for (aeroddc_bank* b : f->banks) FOK(aeroddc_bank_finalize(b));
  f->in_bytes = (size_t)f->B * (f->fmt == AERODDC_CU8 ? 2 : (f->fmt == AERODDC_CS16 ? 4 : 8));
  size_t bytes = 0;
  FOK(aeroddc_bank_host_slot(f->banks[0], 0, &f->h_slot[0], &bytes));
  FOK(aeroddc_bank_host_slot(f->banks[0], 1, &f->h_slot[1], &bytes));
  if (nd > 1) {
    if (!nccl().ok) return aeroddc_set_error(AERODDC_ERR_CUDA, "libnccl.so.2 not found: a fleet of %d GPUs needs NCCL for the raw-block broadcast", nd);
    f->comms.resize(nd);
    FNC(nccl().CommInitAll(f->comms.data(), nd, f->devices.data()));
  }
  f->streams.resize(nd);
  for (int p = 0; p < 2; ++p) { f->d_in[p].resize(nd); f->ev_ready[p].resize(nd); }
  for (int i = 0; i < nd; ++i) {
    FCU(cudaSetDevice(f->devices[i]));
    FCU(cudaStreamCreateWithFlags(&f->streams[i], cudaStreamNonBlocking));
    for (int p = 0; p < 2; ++p) {
      FCU(cudaMalloc((void**)&f->d_in[p][i], f->in_bytes));
      FCU(cudaEventCreateWithFlags(&f->ev_ready[p][i], cudaEventDisableTiming));
    }
  }
  f->finalized = true;
  return AERODDC_OK;
}

int aeroddc_fleet_host_slot(aeroddc_fleet* f, int slot, void** ptr, size_t* bytes) {
  if (!f || !ptr || slot < 0 || slot > 1) return aeroddc_set_error(AERODDC_ERR_ARG, "bad argument");
  if (!f->finalized) return aeroddc_set_error(AERODDC_ERR_STATE, "fleet not finalized");
  *ptr = f->h_slot[slot];
  if (bytes) *bytes = f->in_bytes;
  return AERODDC_OK;
}

int aeroddc_fleet_submit(aeroddc_fleet* f, const void* host_iq, size_t n_complex) {
  if (!f || !host_iq) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL argument");
  if (!f->finalized) return aeroddc_set_error(AERODDC_ERR_STATE, "fleet not finalized");
  if (n_complex != (size_t)f->B) return aeroddc_set_error(AERODDC_ERR_ARG, "block of %zu samples, fleet was created for %d", n_complex, f->B);
  if (f->submitted - f->done >= 2) return aeroddc_set_error(AERODDC_ERR_STATE, "two blocks already in flight; call wait()");
  const int nd = (int)f->banks.size();
  const int p = (int)(f->submitted & 1);
  const void* src = host_iq;
  if (host_iq != f->h_slot[0] && host_iq != f->h_slot[1]) {   // pageable caller memory: stage through the pinned ring
    memcpy(f->h_slot[p], host_iq, f->in_bytes);
    src = f->h_slot[p];
  }
  // the buffers of parity p were last read by the block submitted two calls ago; the in-flight limit above means
  // wait() has retired it (payload copied out, hence kernels done), so no event is needed before overwriting them
  FCU(cudaSetDevice(f->devices[0]));
  FCU(cudaMemcpyAsync(f->d_in[p][0], src, f->in_bytes, cudaMemcpyHostToDevice, f->streams[0]));
  if (nd > 1) {
    FNC(nccl().GroupStart());
    for (int i = 0; i < nd; ++i)
      FNC(nccl().Broadcast(f->d_in[p][0], f->d_in[p][i], f->in_bytes, kNcclUint8, 0, f->comms[i], f->streams[i]));
    FNC(nccl().GroupEnd());
  }
  for (int i = 0; i < nd; ++i) {
    FCU(cudaSetDevice(f->devices[i]));
    FCU(cudaEventRecord(f->ev_ready[p][i], f->streams[i]));
    FOK(aeroddc_bank_submit_device(f->banks[i], f->d_in[p][i], n_complex, f->ev_ready[p][i]));
  }
  f->submitted++;
  return AERODDC_OK;
}

int aeroddc_fleet_wait(aeroddc_fleet* f) {
  if (!f) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL fleet");
  if (f->done >= f->submitted) return aeroddc_set_error(AERODDC_ERR_STATE, "nothing in flight");
  for (aeroddc_bank* b : f->banks) FOK(aeroddc_bank_wait(b));
  f->done++;
  return AERODDC_OK;
}

int aeroddc_fleet_process(aeroddc_fleet* f, const void* host_iq, size_t n_complex) {
  FOK(aeroddc_fleet_submit(f, host_iq, n_complex));
  while (f->done < f->submitted) FOK(aeroddc_fleet_wait(f));
  return AERODDC_OK;
}

int aeroddc_fleet_output(aeroddc_fleet* f, int vfo, const void** payload, size_t* nbytes, uint32_t* rate) {
  if (!f || vfo < 0 || vfo >= (int)f->where.size()) return aeroddc_set_error(AERODDC_ERR_ARG, "vfo index out of range");
  return aeroddc_bank_output(f->banks[f->where[vfo].dev], f->where[vfo].local, payload, nbytes, rate);
}

int aeroddc_fleet_num_devices(aeroddc_fleet* f) { return f ? (int)f->banks.size() : 0; }
int aeroddc_fleet_device_of(aeroddc_fleet* f, int vfo) {
  if (!f || vfo < 0 || vfo >= (int)f->where.size()) return -1;
  return f->where[vfo].dev;
}

int aeroddc_dev_alloc(int device, size_t bytes, void** dev_ptr) {
  if (!dev_ptr) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL argument");
  FCU(cudaSetDevice(device));
  FCU(cudaMalloc(dev_ptr, bytes));
  return AERODDC_OK;
}
int aeroddc_dev_free(int device, void* dev_ptr) {
  FCU(cudaSetDevice(device));
  FCU(cudaFree(dev_ptr));
  return AERODDC_OK;
}
int aeroddc_dev_upload(int device, void* dev_ptr, const void* host, size_t bytes) {
  FCU(cudaSetDevice(device));
  FCU(cudaMemcpy(dev_ptr, host, bytes, cudaMemcpyHostToDevice));
  return AERODDC_OK;
}
int aeroddc_ipc_export(int device, void* dev_ptr, unsigned char handle[AERODDC_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == AERODDC_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  FCU(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  FCU(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle, &h, sizeof h);
  return AERODDC_OK;
}
int aeroddc_ipc_import(int device, const unsigned char handle[AERODDC_IPC_HANDLE_BYTES], void** dev_ptr) {
  if (!dev_ptr) return aeroddc_set_error(AERODDC_ERR_ARG, "NULL argument");
  FCU(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  FCU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return AERODDC_OK;
}
int aeroddc_ipc_close(int device, void* dev_ptr) {
  FCU(cudaSetDevice(device));
  FCU(cudaIpcCloseMemHandle(dev_ptr));
  return AERODDC_OK;
}

int aeroddc_enable_peer(int device, int peer) {
  FCU(cudaSetDevice(device));
  int can = 0;
  FCU(cudaDeviceCanAccessPeer(&can, device, peer));
  if (!can) return aeroddc_set_error(AERODDC_ERR_CUDA, "device %d cannot access device %d", device, peer);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return AERODDC_OK; }
  FCU(e);
  return AERODDC_OK;
}

void aeroddc_fleet_destroy(aeroddc_fleet* f) {
  if (!f) return;
  for (size_t i = 0; i < f->streams.size(); ++i) {
    cudaSetDevice(f->devices[i]);
    if (f->streams[i]) { cudaStreamSynchronize(f->streams[i]); }
  }
  for (aeroddc_bank* b : f->banks) aeroddc_bank_destroy(b);
  for (size_t i = 0; i < f->streams.size(); ++i) {
    cudaSetDevice(f->devices[i]);
    for (int p = 0; p < 2; ++p) {
      if (i < f->d_in[p].size()) cudaFree(f->d_in[p][i]);
      if (i < f->ev_ready[p].size() && f->ev_ready[p][i]) cudaEventDestroy(f->ev_ready[p][i]);
    }
    if (f->streams[i]) cudaStreamDestroy(f->streams[i]);
  }
  for (ncclComm_t c : f->comms) if (c && nccl().ok) nccl().CommDestroy(c);
  delete f;
}

}  // extern "C"
