/*
 * avdsp_oracle.h -- CPU restatement of the AVDSP runtime (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for avdsp_b200.  It is a plain-C restatement of what
 * the reference computes, written from the reference's behaviour (files cited per
 * function in avdsp_oracle.c) with ONE structural change: every piece of state the
 * reference keeps in process globals (sample-rate tables, TPDF/PRNG state,
 * module_avdsp/runtime/dsp_runtime.c:36-38,103-110 and dsp_tpdf.h:11-13,23,33) lives
 * in an instance, so many independent streams can exist in one process.
 *
 * Nothing in the product (avdsp_b200/, include/) may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement
 * bit-for-bit (fmt 2..6) against the reference runtime itself compiled into oracle/_ref
 * (see oracle/Makefile) and against the golden vectors in tests/golden/ that were
 * produced by that compiled reference.  The one exception is fixed-point DSP_FIR,
 * whose reference kernel is provably not a convolution: there the
 * oracle defines the intended semantics and parity for that opcode is "unpinned".
 */
#ifndef AVDSP_ORACLE_H_
#define AVDSP_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct avo_inst avo_inst;

/* Return codes mirror dspRuntimeInit/dspRuntimeReset (dsp_runtime.c:116-195):
 *   >0 totalLength, -1 no header / unknown fs, -2 fs outside header range,
 *   -3 no core, -4 checksum, -5 opcode too recent, -6 program+data > maxWords,
 *   -7 (ours) unsupported DSP_FORMAT / program encoding does not match the format. */
int  avo_create(avo_inst **out, const int32_t *prog, int progWords, int maxWords,
                int format /*2..6*/, int fs, int seed, int defaultDither);
void avo_destroy(avo_inst *);
int  avo_reset(avo_inst *, int fs, int seed, int defaultDither);

int  avo_num_cores(const avo_inst *);
int  avo_total_length(const avo_inst *);
int  avo_data_size(const avo_inst *);
int32_t *avo_code(avo_inst *);            /* private copy of the program words (MEM words live here) */
int32_t *avo_data(avo_inst *);            /* data area, dataSize words */
/* aux state (what the reference holds in globals): [0..3] xoshiro s, [4] tpdfValue, [5] tpdfRandom,
 * [6] current global dither, [7] default dither */
void avo_get_aux(const avo_inst *, int32_t aux[8]);
void avo_set_aux(avo_inst *, const int32_t aux[8]);

/* One core for one frame == dspRuntime_<fmt>(corePtr, data, io) (dsp_runtime.c:302). core is 1-based. */
int  avo_run_core(avo_inst *, int core, int32_t *io);
/* Canonical order: one frame, cores ascending, one shared io[] */
int  avo_run_frame(avo_inst *, int32_t *io);

/* Batched canonical processing of interleaved PCM: in[nFrames][nIn], out[nFrames][nOut].
 * io[] is zeroed at the start of every frame, inputs placed at inIdx[], outputs read from outIdx[].
 * For formats 5/6 samples are IEEE float bit patterns carried in int32. */
int  avo_process(avo_inst *, const int32_t *in, int32_t *out, int nFrames,
                 const int *inIdx, int nIn, const int *outIdx, int nOut);
/* Emulation of the ALSA plugin loop nest (linux/avdsp_plugin.c:95-142): core-major within
 * each period, fresh io[] per (core,frame) holding only that core's used inputs. */
int  avo_process_plugin_order(avo_inst *, const int32_t *in, int32_t *out, int nFrames, int period,
                              int nIn, int nOut);

#ifdef __cplusplus
}
#endif
#endif
