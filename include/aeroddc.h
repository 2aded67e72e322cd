/* aeroddc.h - C ABI of the B200-native multi-VFO digital down-converter bank.
 *
 * The reference (airframesio/aero-cli, aero-publish) has no plugin/FFI seam for this path: the
 * seam is the C++ class `vfo` and the .ini semantics of Publisher::loadSettings. Each entry point
 * below names the reference interface it stands in for (paths under /root/reference/publish).
 * The C++ classes in aero-cli_b200/host/ (`vfo`, `Publisher`) keep the reference's method names on
 * top of this ABI; INTEGRATION.md shows the few lines a maintainer changes in publisher.cpp.
 *
 * Conventions: plain pointers and sizes; every function returns AERODDC_OK (0) or a negative
 * error code and records a message retrievable with aeroddc_last_error(); no exceptions cross the
 * boundary; a bank handle is NOT thread-safe (the reference drives all VFOs from one worker
 * thread, publisher.cpp:229-232,301-305). There is no CPU fallback: without a CUDA device every
 * compute entry point fails with AERODDC_ERR_CUDA.
 */
#ifndef AERODDC_H
#define AERODDC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AERODDC_ABI_VERSION 1

enum {
  AERODDC_OK = 0,
  AERODDC_ERR_ARG = -1,    /* bad argument / contract violation (block size, divisibility, ...) */
  AERODDC_ERR_STATE = -2,  /* call order (add after finalize, process before finalize, ...)     */
  AERODDC_ERR_CUDA = -3,   /* CUDA runtime error, or no device                                  */
  AERODDC_ERR_DESIGN = -4, /* filter design rejected the parameters (firfilter.cpp:100-112)     */
  AERODDC_ERR_NOMEM = -5
};

/* Raw IQ sample formats of a block (interleaved I,Q). cf32 is what the reference receives from
 * SoapySDR (publisher.cpp:254,267-268); cu8 and cs16 are the IQ-file-source formats, converted as
 * (u8 - 127.4f) / 128.0f and s16 / 32768.0f. */
enum { AERODDC_CU8 = 0, AERODDC_CS16 = 1, AERODDC_CF32 = 2 };

typedef struct aeroddc_bank aeroddc_bank;

/* One VFO, i.e. the arguments of the vfo setters + init (vfo.h:16-40, vfo.cpp:57-139) as
 * Publisher::loadSettings fills them (publisher.cpp:118-148,159-219). */
typedef struct aeroddc_vfo_desc {
  double mixer_freq;   /* vfo::setMixerFreq: Hz, NCO frequency relative to the bank's centre          */
  int decim_count;     /* vfo::setDecimationCount: number of half-band /2 stages D, 0..8             */
  int late_decimate;   /* vfo::init(.., lateDecimate): 0, or 5/6 for the final FIR decimation        */
  int filter_bw;       /* vfo::setFilterBandwidth: Hz, 0 = no fir_usb                                */
  float gain;          /* vfo::setGain (the ini value / 100, publisher.cpp:213)                      */
  int demod_usb;       /* vfo::setDemodUSB: 1 = USB-demodulated int16 audio, 0 = compressed IQ       */
  int compress_style;  /* vfo::setCompressonStyle: 1 = nibble-packed, 0 = int8 I,Q (demod_usb == 0)  */
  int scale_comp;      /* vfo::setScaleComp (>0)                                                     */
  char topic[64];      /* vfo::setZmqTopic; the wire carries its first 5 bytes (zmqpublisher.cpp:69) */
  int parent;          /* -1: fed by the bank's raw IQ. Otherwise the index of an earlier VFO (a "main" VFO,
                          publisher.cpp:118-148) whose stage-D stream is this VFO's input, as vfo::setVFOs
                          + the recursion in vfo::process do (vfo.cpp:167-172). A VFO that has children
                          publishes nothing itself, like the reference. Its children run at its output
                          rate Fs/2^D with blocks of block_len/2^D samples (publisher.cpp:215-217). */
} aeroddc_vfo_desc;

/* Create an empty bank for blocks of exactly `block_len` complex samples at rate `sample_rate`
 * in format `in_format`, on CUDA device `device`.
 * Replaces: the per-VFO setFs()/init(samplesPerBuffer) pair (vfo.cpp:57-139, 141-142) and the
 * buflen arithmetic of Publisher::loadSettings (publisher.cpp:93-100). */
int aeroddc_bank_create(aeroddc_bank **out, int sample_rate, int block_len, int in_format, int device);

/* How the bank cuts a block of one VFO group (VFOs sharing input stream and DA = min(D, 5), the half-band stages the
 * main kernel runs in registers; stages 5..D-1 run in the deep kernel at <= 1/32 of the rate and need no plan) into
 * work for the main kernel (pure host arithmetic, no device needed; exposed for inspection and tests). All lengths in
 * input samples. Invariants: segment_len and part_len are multiples of 256; n_segments*segment_len >= block_len;
 * parts*part_len >= segment_len; segment_len >= 2*warmup. */
typedef struct aeroddc_segment_plan {
  int warmup;           /* W = 10*2^DA (rounded to the chunk): samples a segment re-processes to rebuild its history */
  int boundary_warmup;  /* 11*2^D: samples the boundary CTA re-processes for the next block's shifted history      */
  int segment_len, n_segments;
  int parts, part_len;  /* a full segment runs as `parts` chained CTAs of part_len samples (the last one may need fewer) */
  int vfo_groups;       /* CTAs side by side: ceil(n_vfos / 32), one warp of 32 VFOs each                          */
  int ctas;             /* grid size of the launch                                                                 */
} aeroddc_segment_plan;
int aeroddc_plan_segments(int block_len, int decim_count, int n_vfos, int n_sm, double waves, int parts, aeroddc_segment_plan *out);

/* Host-side planning of the tensor mode (below): the (64-VFO tile, stage-5 output) plane of n_tiles x n_mid outputs is cut
 * into one contiguous run per persistent CTA. With cta_first == stretches == NULL returns the CTA count P (<= n_sm).
 * Otherwise fills cta_first[P + 1] (stretches of CTA c: [cta_first[c], cta_first[c + 1])) and stretches[3 * i] =
 * {tile, first output, end output} (first outputs are 0 or multiples of 256) and returns the number of stretches.
 * Pure host code; replaces nothing in the reference (its loop is per VFO, vfo.cpp:154-186). */
int aeroddc_plan_tensor_stretches(int n_tiles, int n_mid, int n_sm, int *cta_first, int *stretches, int cap);

/* Arithmetic mode of the half-band/mix/NCO kernel; call before the first block.
 *   AERODDC_MODE_EXACT (default): every multiply and add of the reference, un-fused, in its order:
 *       payloads are byte-identical to the reference's vfo::process chain.
 *   AERODDC_MODE_FAST: the same chain with fused multiply-adds and a rotation-only oscillator between
 *       exact checkpoints; NOT bit-identical, within BASELINE.json's tolerance (max |err| <= 1e-4 of
 *       full scale, error SNR >= 80 dB at normal signal levels), ~1.4x the throughput.
 *   AERODDC_MODE_TENSOR: the mix and the first five half-band stages as one complex GEMM on the 5th-generation
 *       tensor cores (tcgen05, bf16 hi+mid operand split, fp32 accumulation in TMEM; csrc/tc_kernels.cuh); same
 *       tolerance class as AERODDC_MODE_FAST. Applies to VFOs with decim_count >= 5 fed by the bank's raw cf32
 *       stream (or any format with DC correction on); every other VFO, the head of each block
 *       and the zone after an oscillator-table restart run as in AERODDC_MODE_FAST. Set before finalize. */
enum { AERODDC_MODE_EXACT = 0, AERODDC_MODE_FAST = 1, AERODDC_MODE_TENSOR = 2 };
int aeroddc_bank_set_mode(aeroddc_bank *bank, int mode);

/* Optional DC removal on the raw stream before any VFO, as Publisher::demodData does with --enable-dcc /
 * correct_dc_bias=1 (publisher.cpp:292-296): avept = avept*(1-1e-6f) + 1e-6f*x; x -= avept, per rail, float32.
 * The recurrence is sequential and is executed as such (one GPU thread per rail, ~8 cycles per sample: negligible
 * at the reference's rates, about 60 ms per block at 61.44 MS/s). Call before finalize. */
int aeroddc_bank_set_dc_correction(aeroddc_bank *bank, int enable);

/* Append a VFO; returns its index (>= 0) or a negative error code.
 * Replaces: `new vfo()` + setters in publisher.cpp:121-147 and :159-217. */
int aeroddc_bank_add_vfo(aeroddc_bank *bank, const aeroddc_vfo_desc *desc);

/* Design all filters on the host (firfilter::low_pass, FIRHilbert taps, oscillator rotation;
 * firfilter.cpp:46-99, dsp.cpp:181-215, oscillator.cpp:8-10), upload them, and run the NCO
 * checkpoint kernel (Oscillator::Oscillator, oscillator.cpp:12-27). Replaces: vfo::init. */
int aeroddc_bank_finalize(aeroddc_bank *bank);

/* Process one block given in HOST memory: host->device copy, all kernels, device->host copy of
 * every VFO's payload, then returns (synchronous). n_complex must equal block_len (the reference's
 * block contract, vfo.cpp:155,164).
 * Replaces: Publisher::demodData -> vfo::process for every VFO (publisher.cpp:285-306,
 * vfo.cpp:154-186), up to but excluding ZmqPublisher::publish. */
int aeroddc_bank_process(aeroddc_bank *bank, const void *host_iq, size_t n_complex);

/* Pipelined form of the same: submit() enqueues H2D + kernels + D2H for one block and returns;
 * wait() blocks until the OLDEST submitted block's payloads are in host memory and makes them
 * the ones aeroddc_bank_output() returns. At most two blocks may be in flight. host_iq must stay
 * valid until the matching wait() unless it lies in the bank's own pinned ring (next call). */
int aeroddc_bank_submit(aeroddc_bank *bank, const void *host_iq, size_t n_complex);
int aeroddc_bank_wait(aeroddc_bank *bank);

/* Rewind the bank to stream position 0: all filter history, the oscillator position, the stage-D history and the DC
 * average return to their freshly initialised values (as after vfo::init, vfo.cpp:57-139), the VFO set and every
 * design table stay. Nothing may be in flight. Used when a file source is re-opened and by bench.py's parity leg,
 * which replays six known blocks through the very bank object it has just timed. */
int aeroddc_bank_reset(aeroddc_bank *bank);

/* Pinned host staging ring ("pinned host ring plus async H2D", replaces the malloc'ed samplesBuf
 * of Publisher::readerThread, publisher.cpp:238,267-268): slot 0/1, each block_len samples. A
 * source that fills these directly avoids one host copy. */
int aeroddc_bank_host_slot(aeroddc_bank *bank, int slot, void **ptr, size_t *bytes);

/* Same chain on a block that already lies in DEVICE memory (e.g. the destination of an NCCL
 * broadcast): enqueues the kernels and the payload D2H on the bank's stream after `ready_event`
 * (a cudaEvent_t, may be NULL) and returns. Finish with aeroddc_bank_wait(). */
int aeroddc_bank_submit_device(aeroddc_bank *bank, const void *dev_iq, size_t n_complex, void *ready_event);

/* The same with the block spread over n_slices device buffers, slice i holding samples [i*slice_len, (i+1)*slice_len)
 * (the last one may be shorter). The buffers may live in the HBM of DIFFERENT GPUs (peer-mapped addresses: CUDA IPC
 * or cudaDeviceEnablePeerAccess): the main kernel's TMA tile loads then pull every tile from the GPU that holds it,
 * over NVLink, while computing - the multi-GPU exchange of the raw block fused into the
 * kernel, with no broadcast step and no staging copy; when every GPU of a node ingests 1/N of each block over its own
 * PCIe link, all N links and all N NVLink ports share the load. A bank in AERODDC_MODE_TENSOR, which would pull every
 * tile across NVLink once per 64-VFO tile and outrun the links, instead lets its copy engines gather the slices into its
 * own input buffer first (one block ahead on its copy stream; each byte crosses once) and computes from local memory.
 * slice_len must be a multiple of 32 and >= 256; n_slices <= 8. The kernels start after every event of ready_events[0..n_events) (cudaEvent_t, NULL entries skipped). */
int aeroddc_bank_submit_device_sliced(aeroddc_bank *bank, const void *const *slices, int n_slices, size_t slice_len,
                                      size_t n_complex, void *const *ready_events, int n_events);

/* Payload of VFO `vfo` for the most recently completed block: pointer into pinned host memory
 * (valid until the next wait()/process()), its length in bytes and the output sample rate.
 * Replaces: transmit_usb / transmit_iq and outputRate as handed to ZmqPublisher::publish
 * (vfo.cpp:289-313, zmqpublisher.cpp:61-73): frame 3 and frame 2 of the ZeroMQ message. */
int aeroddc_bank_output(aeroddc_bank *bank, int vfo, const void **payload, size_t *nbytes, uint32_t *rate);

/* Topic string of a VFO (frame 1 is its first 5 bytes). */
const char *aeroddc_bank_topic(aeroddc_bank *bank, int vfo);

/* Copy the stage-D complex stream (vfo::decimate[decimateCount], vfo.h:39) of the most recently
 * completed block to host memory as interleaved float I,Q; returns the number of complex samples
 * or a negative error. Test/debug hook for stage-wise parity. */
int aeroddc_bank_stage_d(aeroddc_bank *bank, int vfo, float *host_out, size_t cap_complex);

int aeroddc_bank_num_vfos(aeroddc_bank *bank);

/* Device time, in milliseconds, spent by the kernels of the most recently completed block
 * (CUDA events on the bank's compute stream) and the number of kernels launched for it. */
int aeroddc_bank_last_timing(aeroddc_bank *bank, float *kernel_ms, int *launches);

/* Device time of the dominant kernel (the fused unpack + mix + half-band cascade) alone. */
int aeroddc_bank_last_main_ms(aeroddc_bank *bank, float *main_ms);

/* Stopwatch on the bank's own streams, for measuring several blocks as one region with CUDA events:
 * which = 0 records the start event on the compute stream (inputs already in HBM),
 * which = 1 records the start event on the copy stream (host inputs; the H2D copies are inside),
 * which = 2 records the stop event on the compute stream, waits for it, and returns the elapsed
 * milliseconds since the last start in *ms. */
int aeroddc_bank_stopwatch(aeroddc_bank *bank, int which, float *ms);

/* Resource use after finalize: bytes of HBM held by the bank. */
int aeroddc_bank_device_bytes(aeroddc_bank *bank, size_t *bytes);

void aeroddc_bank_destroy(aeroddc_bank *bank);

/* Last error message of this thread ("" if none). */
const char *aeroddc_last_error(void);

/* Measure the FP32 FFMA issue peak of `device` with a register-only microbenchmark (the roofline
 * denominator SURVEY.md section 8d asks to measure in the same run): TFLOP/s counting FMA = 2. */
int aeroddc_measure_fp32_peak(int device, double *tflops, double *sm_clock_mhz);

/* ------------------------------------------------------------------------------------------------
 * Fleet: one bank per GPU of a node, driven from ONE host thread.
 * VFOs are sharded over the devices (flat VFO i -> device i mod N; a main VFO takes its sub-VFOs with it), each GPU
 * runs its VFO subset and returns its own payloads. The raw block reaches the GPUs in one of two ways:
 *   peer (default): GPU i uploads slice i (1/N of the block) over its own PCIe link - N concurrent H2D copies - and
 *     every bank's kernel reads all slices in place through peer memory over NVLink (aeroddc_bank_submit_device_sliced;
 *     in AERODDC_MODE_TENSOR the bank's copy engines gather them into its own input buffer first, see there);
 *   nccl (AERODDC_EXCHANGE=nccl, or no peer access): the block is uploaded once to devices[0] and broadcast to the
 *     others with ncclBroadcast over NVLink (libnccl.so.2 is loaded at run time).
 * Same call order and error conventions as the bank; VFO indices are global (order of add_vfo).
 * Replaces, like the bank, Publisher::demodData -> vfo::process for every VFO (publisher.cpp:285-306).
 * ---------------------------------------------------------------------------------------------- */
typedef struct aeroddc_fleet aeroddc_fleet;
int aeroddc_fleet_create(aeroddc_fleet **out, int sample_rate, int block_len, int in_format, const int *devices, int n_devices);
int aeroddc_fleet_add_vfo(aeroddc_fleet *fleet, const aeroddc_vfo_desc *desc);   /* desc->parent is a global index */
int aeroddc_fleet_set_mode(aeroddc_fleet *fleet, int mode);
int aeroddc_fleet_set_dc_correction(aeroddc_fleet *fleet, int enable);
int aeroddc_fleet_finalize(aeroddc_fleet *fleet);
/* Pinned host slot 0/1 of the ingest GPU's ring (fill it directly to avoid a host copy). */
int aeroddc_fleet_host_slot(aeroddc_fleet *fleet, int slot, void **ptr, size_t *bytes);
/* submit: the uploads (or upload + broadcast) + every GPU's kernels and payload D2H, asynchronously (at most two
 * blocks in flight); wait: the oldest submitted block's payloads are in host memory on every GPU. wait retires the
 * block on every GPU even when one reports an error; after any error the fleet refuses further blocks. */
int aeroddc_fleet_submit(aeroddc_fleet *fleet, const void *host_iq, size_t n_complex);
int aeroddc_fleet_wait(aeroddc_fleet *fleet);
int aeroddc_fleet_process(aeroddc_fleet *fleet, const void *host_iq, size_t n_complex);
int aeroddc_fleet_reset(aeroddc_fleet *fleet);   /* aeroddc_bank_reset on every GPU */
int aeroddc_fleet_output(aeroddc_fleet *fleet, int vfo, const void **payload, size_t *nbytes, uint32_t *rate);
int aeroddc_fleet_num_devices(aeroddc_fleet *fleet);
int aeroddc_fleet_exchange(aeroddc_fleet *fleet);             /* 0 = single GPU, 1 = peer slices, 2 = NCCL broadcast */
int aeroddc_fleet_device_of(aeroddc_fleet *fleet, int vfo);   /* index into the devices[] given at create */
void aeroddc_fleet_destroy(aeroddc_fleet *fleet);

/* ------------------------------------------------------------------------------------------------
 * Peer-memory exchange for one-process-per-GPU deployments. aeroddc_bank_submit_device accepts ANY device
 * address the GPU can read - including a block that lives in another GPU's HBM: the main kernel's TMA tile
 * loads then pull the raw samples across NVLink tile by tile while it computes, and no separate broadcast
 * step or staging copy exists. These helpers create such a block in one process and map it in the others
 * (CUDA IPC): export on the owner, send the 64-byte handle through any channel, import on the peers.
 * ---------------------------------------------------------------------------------------------- */
#define AERODDC_IPC_HANDLE_BYTES 64
int aeroddc_dev_alloc(int device, size_t bytes, void **dev_ptr);
int aeroddc_dev_free(int device, void *dev_ptr);
int aeroddc_dev_upload(int device, void *dev_ptr, const void *host, size_t bytes);   /* synchronous H2D */
int aeroddc_dev_upload_async(int device, void *dev_ptr, const void *pinned_host, size_t bytes, void *stream);   /* cudaMemcpyAsync on a cudaStream_t */
int aeroddc_ipc_export(int device, void *dev_ptr, unsigned char handle[AERODDC_IPC_HANDLE_BYTES]);
int aeroddc_ipc_import(int device, const unsigned char handle[AERODDC_IPC_HANDLE_BYTES], void **dev_ptr);
int aeroddc_ipc_close(int device, void *dev_ptr);
/* Same process, two devices: let `device` read `peer`'s memory (cudaDeviceEnablePeerAccess; idempotent). */
int aeroddc_enable_peer(int device, int peer);

/* Host-side coefficient designers, exposed so that callers and tests can inspect exactly the taps
 * the bank uploads. Pure CPU code (no device needed), bit-identical to firfilter::low_pass with the
 * Hamming window (firfilter.cpp:46-99,186-193), FIRHilbert::FIRHilbert (dsp.cpp:181-215) and the
 * oscillator rotation (oscillator.cpp:8-10). The tap functions return the tap count (writing at
 * most `cap` taps) or AERODDC_ERR_DESIGN when the reference would throw (firfilter.cpp:100-112). */
int aeroddc_design_lowpass(double gain, double fs, double cutoff, double transition, float *taps, int cap);
int aeroddc_design_hilbert(int len, int fs_param, float *taps, int cap);
int aeroddc_design_rotation(double fs, double freq, float *cos_out, float *sin_out);

int aeroddc_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AERODDC_H */
