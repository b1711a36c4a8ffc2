/*
 * trialign.h -- C ABI of libtrialign, the B200-native engine for the three-sequence alignment score
 * that the reference RTL (TRIALIGN / PE / MAX7) computes.
 *
 * The reference has no host-language layer and no FFI (its .gitignore withholds the host scripts); the only
 * interface its hot path sits behind is the TRIALIGN module port list, driven by the testbench:
 *
 *   reference interface                                      replaced here by
 *   -------------------------------------------------------  ------------------------------------------
 *   TRIALIGN ports  A_idx/B_idx/C_idx, A/B/C_symbol,          trialign_score(): lengths + symbol arrays in,
 *     A/B/C_addr, start_align, finish, Score                  Score out (signed, SCORE_BITS wide)
 *     (src/TriAlign_1cyc.v:9-22, src/TriAlign_tb.sv:204-262)
 *   parameter SCORE_BITS (src/TriAlign_1cyc.v:6)              the score_bits argument (run-time)
 *   parameter PE_LEN / SRAM_ADDR_BITS / *_TOTAL_LEN           none: array size, face depth and slicing are
 *     (src/TriAlign_1cyc.v:1-8)                               internal to the CUDA kernels
 *   dat/A_seq.dat, B_seq.dat, C_seq.dat (one digit per line)   trialign_read_seq_dat()
 *   `$display("TriAlign Score: %d")` (TriAlign_tb.sv:339-341)  trialign_write_scores(): "%d" per line,
 *     and dat/Pattern1.dat's one-integer-per-line layout       '\n'-joined, no trailing newline
 *   tb symbol codes A=0,T=1,C=2,G=3,N=4 (TriAlign_tb.sv:42-46) bytes; only bits [1:0] are used
 *     and the PE's 2-bit truncation (src/PE_1cyc.v:63-66)       (N aliases to A exactly as in the RTL)
 *
 * Semantics (bit-exact with the RTL): 7 states per cell (M,Ix,Iy,Iz,Ixy,Iyz,Ixz), scoring constants of
 * src/PE_1cyc.v:55-61, substitution scores of :159-162 (including the operator-precedence quirk of temp_ABC),
 * all-zero boundary planes (src/TriAlign_1cyc.v:156-181), every candidate wrapped to score_bits two's complement
 * before the signed max (src/PE_1cyc.v:127-133), Score = max of the 7 states of cell (la,lb,lc).
 *
 * Conventions: C linkage, plain pointers and sizes, no exceptions cross the boundary, every function returns
 * TRIALIGN_OK (0) or a negative error and never aborts.  Host-pointer entry points copy to the GPU and back
 * themselves; *_dev entry points take device pointers and enqueue on the given CUDA stream.  There is no CPU
 * fallback: without a usable CUDA device every compute entry point returns TRIALIGN_ENODEVICE.
 */
#ifndef TRIALIGN_H
#define TRIALIGN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRIALIGN_VERSION 0x00010000 /* major<<16 | minor<<8 | patch */

enum trialign_rc {
    TRIALIGN_OK = 0,
    TRIALIGN_EINVAL = -1,       /* bad pointer, negative length, score_bits outside [2,32], ... */
    TRIALIGN_ENOMEM = -2,       /* host or device allocation failed */
    TRIALIGN_ECUDA = -3,        /* a CUDA call failed; see trialign_last_cuda_error() */
    TRIALIGN_EIO = -4,          /* file could not be opened / read / written */
    TRIALIGN_EPARSE = -5,       /* dat file holds something other than digits 0-3 and line ends */
    TRIALIGN_EUNSUPPORTED = -6, /* backend not part of the product (CPU checkers live under oracle/), or a size beyond a
                                   kernel's limits (e.g. more than 8,192 planes in one slab launch) */
    TRIALIGN_ENODEVICE = -7,    /* no usable CUDA device */
    TRIALIGN_ECAPACITY = -8,    /* output buffer too small */
    TRIALIGN_EINEXACT = -9      /* TRIALIGN_BACKEND_GPU_FAST forced, but wrap-freedom could not be established */
};

enum trialign_backend {
    TRIALIGN_BACKEND_AUTO = 0,        /* GPU: fast kernels where provably bit-exact, literal kernel otherwise */
    TRIALIGN_BACKEND_CPU_ORACLE = 1,  /* reserved numbers of the survey's boundary; not in the product */
    TRIALIGN_BACKEND_RTL_1CYC = 2,
    TRIALIGN_BACKEND_RTL_2CYC = 3,
    TRIALIGN_BACKEND_GPU = 4,         /* same as AUTO */
    TRIALIGN_BACKEND_GPU_LITERAL = 5, /* force the literal 49-add/42-max kernel (any score_bits in [2,32]) */
    TRIALIGN_BACKEND_GPU_FAST = 6     /* force the factored packed-16-bit kernels; EINEXACT if a wrap was possible */
};

/* Which kernel family served the last call on this thread (trialign_last_path). */
enum trialign_path {
    TRIALIGN_PATH_NONE = 0,
    TRIALIGN_PATH_LITERAL = 1,        /* plane-sweep literal kernel */
    TRIALIGN_PATH_FAST_BATCH = 2,     /* K1: batched systolic kernel, two triples per 32-bit lane */
    TRIALIGN_PATH_FAST_LONG = 3,      /* K2: block-dataflow kernel for one long triple */
    TRIALIGN_PATH_FAST_THEN_LITERAL = 4, /* range-checked fast kernel ran; the triples whose values came near the
                                            score_bits limits (all of a single-triple call, only the flagged ones of a
                                            batch: trialign_last_flagged) were re-scored by the literal kernel */
    TRIALIGN_PATH_TRACEBACK = 5,      /* KT: literal plane sweep with a direction cube + backward walk */
    TRIALIGN_PATH_FAST_STREAM = 6     /* K2S: streaming block-sweep kernel for batches of triples longer than 64 */
};

int trialign_version(void);
/* Scoring parameters.  The RTL hard-wires them as PE localparams (src/PE_1cyc.v:53-59: MATCH=1,
 * MISMATCH=-1, GO=2, GE=1; GO2 = GO<<1, GE2 = GE<<1, GOGE = GO+GE derive from them) and its testbench carries the intent
 * to make them inputs: the commented-out ports .i_match / .i_mismatch / .i_gap_extend / .i_gap_open
 * (src/TriAlign_tb.sv:220-224).  The *_ex entry points take them as arguments. */
enum trialign_triple_rule {
    TRIALIGN_TRIPLE_RTL = 0,          /* temp_ABC exactly as src/PE_1cyc.v:162 parses: 3*MATCH if a=b=c,
                                         (MATCH+MISMATCH)<<1 if a=b!=c, 3*MISMATCH whenever a!=b */
    TRIALIGN_TRIPLE_SUM_OF_PAIRS = 1  /* the paper's S(a,b,c) = s(a,b)+s(b,c)+s(a,c) (pic/Algorithm.png) */
};
typedef struct trialign_scoring {
    int32_t match;       /* MATCH: score of an equal pair */
    int32_t mismatch;    /* MISMATCH: score of an unequal pair (signed: the RTL's is -1) */
    int32_t gap_open;    /* GO: cost of opening a gap in one sequence (>= 0) */
    int32_t gap_extend;  /* GE: cost of extending it (>= 0) */
    int32_t triple_rule; /* enum trialign_triple_rule */
} trialign_scoring;

const char *trialign_strerror(int rc);
const char *trialign_last_cuda_error(void); /* thread-local text of the last CUDA failure, "" if none */
int trialign_device_count(void);            /* 0 when no driver / no device; never negative */

/* Run-time options: the experiment switches of the kernels (A/B variants, gate intervals, multi-GPU decomposition).
 * Nothing in the library reads the environment; a caller that wants a non-default kernel says so explicitly.
 * Names: debug, k1, k2s, k2_lag, k2_pp, k2_gatetid, k2_early, k2_gate, k2_rgate, k2_cols_order, multi ("slab" | "cols";
 * default: chosen by trialign_multi_plan), multi_min, kt_seg, kt_budget_mb (traceback, see trialign_align).
 * value NULL or "" restores the default.  Unknown names give TRIALIGN_EINVAL.  Process-wide, thread-safe. */
int trialign_set_option(const char *name, const char *value);
int trialign_get_option(const char *name, char *buf, int cap); /* "" when unset */

/* One triple.  a/b/c: host bytes.  score_bits in [2,32].  devices: bitmask of CUDA ordinals (0 = device 0); a bit
 * beyond trialign_device_count() gives TRIALIGN_EINVAL.
 * final_states (may be NULL) receives the 7 states of cell (la,lb,lc) in the order M,Ix,Iy,Iz,Ixy,Iyz,Ixz.
 * Any zero length gives score 0 (the cell lies on a zero boundary plane).
 * Several bits set: the calling thread drives all of those GPUs (peer access, one stream per device) and the triple is
 * split as an i-slab pipeline -- GPU g computes planes (la*g/G, la*(g+1)/G] and hands the plane state of every block to
 * GPU g+1 through NVLink peer stores -- or, from 4 GPUs on, as a column pipeline (GPU g sweeps all planes of a range of
 * block columns and streams the boundary faces to GPU g+1; trialign_multi_plan tells which) -- provided every GPU gets
 * at least 64 planes / 128 columns and the fast kernels are provably bit-exact for the input; otherwise the first GPU
 * of the mask serves the call alone.  trialign_last_devices() reports how many GPUs did. */
int trialign_score(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                   int backend, unsigned devices, int32_t *score, int32_t final_states[7]);

/* The same with explicit scoring parameters.  trialign_scoring_rtl() fills in the RTL's constants; with exactly those
 * the call is identical to trialign_score / trialign_score_batch (all kernels).  With other values a BATCH of triples of
 * up to 2,048 symbols runs the factored kernels with the constants as a kernel argument, range-checked on every step
 * and with the flagged triples re-scored literally, provided 0 <= GE <= 4, GE <= GO <= 8, -8 <= MISMATCH <= MATCH <= 8,
 * MISMATCH >= -2 GE and score_bits >= 6; single triples and everything else run the literal kernel (every candidate
 * wrapped to score_bits before the signed max, as the RTL's datapath would with other localparams).  The result does
 * not depend on the kernel.  |match|, |mismatch|, gap_open, gap_extend <= 1024, gap costs >= 0, else TRIALIGN_EINVAL. */
void trialign_scoring_rtl(trialign_scoring *sc);
int trialign_score_ex(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                      const trialign_scoring *scoring, unsigned devices, int32_t *score, int32_t final_states[7]);
int trialign_score_batch_ex(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len, int score_bits,
                            const trialign_scoring *scoring, unsigned devices, int32_t *scores);

/* Uniform-length batch, structure of arrays: a/b/c are [n][len] row-major host bytes, scores is [n].
 * With several bits set in `devices` the batch is split contiguously across those GPUs (no exchange). */
int trialign_score_batch(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len, int score_bits,
                         int backend, unsigned devices, int32_t *scores);

/* Device-resident variants: every pointer is device memory on CUDA ordinal `device`, work is enqueued on
 * `stream` (a cudaStream_t; NULL = legacy default stream) and the call returns without synchronising unless
 * the range check of the fast path forces a literal re-run.  d_scores is [n] int32.
 * The calling thread's current CUDA device is left unchanged.  Calls on one device share that device's scratch
 * buffers (block faces, progress words): a call enqueued on another stream than its predecessor first waits (on the
 * device, cudaStreamWaitEvent) for the predecessor's kernels, so calls in flight on different streams are safe but run
 * one after the other. */
int trialign_score_batch_dev(int device, const uint8_t *d_a, const uint8_t *d_b, const uint8_t *d_c, int64_t n,
                             int len, int score_bits, int backend, int32_t *d_scores, void *stream);
/* d_out: int32[8] = score followed by the 7 final states. */
int trialign_score_dev(int device, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
                       int lc, int score_bits, int backend, int32_t *d_out, void *stream);

/* Optimal alignment of one triple (host pointers): the traceback of the recurrence, defined as what the MAX7
 * instances' commented-out `act` outputs would select (src/PE_1cyc.v:12,22-30,139-145; SURVEY.md 8(f).2 -- the RTL
 * itself ships score-only).  The path runs from the border cell start[] = (i0,j0,k0) (one of them 0: all three border
 * planes are zero, src/TriAlign_1cyc.v:156-181, so the symbols before it are unaligned and free) to (la,lb,lc).
 * The traceback runs on ONE GPU: the first of the `devices` mask (the direction cube is not partitioned).
 * moves[0..*n_moves): one byte per alignment column, first column first: bit0 = A advances, bit1 = B, bit2 = C
 * (7 = three symbols aligned / state M, 1 = Ix, 2 = Iy, 4 = Iz, 3 = Ixy, 6 = Iyz, 5 = Ixz).  *score equals
 * trialign_score()'s.  cap < *n_moves gives ECAPACITY (cap = la+lb+lc always suffices).  Any score_bits in [2,32].
 * Memory: the direction cube takes la*lb*lc*4 bytes; when that exceeds the budget (option kt_budget_mb, default 85 %
 * of the free device memory) the sweep is cut into segments of D anti-diagonal planes with the rolling planes
 * checkpointed at the boundaries, and every segment but the last is swept a second time when the walk reaches it:
 * D*lb*lc*4 + ((la+lb+lc)/D - 2) * 28*(lb+1)*(lc+1)*sizeof(score) bytes, D ~ sqrt(7*(la+lb+lc)*sizeof(score)) (option
 * kt_seg forces D).  ENOMEM only when even that does not fit.  trialign_last_segments() tells which form ran. */
int trialign_align(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                   unsigned devices, int32_t *score, int32_t start[3], uint8_t *moves, int cap, int *n_moves);

/* Every triple of a sequence set (the progressive three-way alignment of pic/Motivation.png scores
 * all triples of its input -- the workload behind "a million short triples": 186 sequences give 1,055,240).
 * seqs: [nseq][len] symbols, uploaded once; the C(nseq,3) triples are laid out on the device and scored by the batch
 * kernels.  scores[t] belongs to the t-th triple (x < y < z) in lexicographic order with roles A = x, B = y, C = z
 * (the RTL's score is not symmetric in the roles, src/PE_1cyc.v:162); trialign_triple_count(nseq) = C(nseq,3) entries.
 * With several bits set in `devices` the triple list is split contiguously across those GPUs (no exchange). */
int64_t trialign_triple_count(int nseq);
int trialign_score_all_triples(const uint8_t *seqs, int nseq, int len, int score_bits, int backend, unsigned devices,
                               int32_t *scores);

/* How trialign_score() would spread one long triple over n_devices GPUs: 0 = first GPU alone (too short, or not provably
 * wrap-free), 1 = i-slab pipeline, 2 = column pipeline (option multi = slab | cols forces one; the default picks columns
 * from 4 GPUs on when every GPU gets at least two 128-column block pairs).  One process per GPU (below) should follow
 * the same plan; parallel.make_pipeline() does. */
int trialign_multi_plan(int la, int lb, int lc, int score_bits, int n_devices);

/* ---- i-slab pipeline: ONE long triple over several GPUs, one process per GPU ----
 * Rank g computes planes (i_begin, i_end] of the cube for every (j,k) block and hands the plane-i_end state of each
 * block task to rank g+1, whose kernel picks it up task by task while rank g is already working on later blocks.  The
 * hand-over buffers (state + one int flag per block task) live on the RECEIVING GPU; the sender writes them through a
 * peer mapping (NVLink), so there is no host round trip and no collective.  Sizes come from
 * trialign_slab_buffer_bytes(); in_* = this rank's own buffers (NULL on the first rank), out_* = the next rank's
 * buffers as mapped with trialign_ipc_open() (NULL on the last rank).  `epoch` >= 1 must differ between consecutive runs
 * that reuse the buffers (flags are compared for equality, never cleared), and -- there is ONE hand-over slot per block
 * task -- run r+1 may only be enqueued on rank g once run r has finished on rank g+1 (parallel.SlabPipeline.run() orders
 * the runs with a stream-ordered collective; trialign_score() with a device mask synchronises every GPU before it
 * returns).  Only the last rank writes d_out (int32[8]: score + 7 final states).  Requires an input that is provably
 * wrap-free at score_bits (else EINEXACT); a slab of more than 8,192 planes, or a triple whose values leave the 16-bit
 * potential range, gives EUNSUPPORTED (la itself may exceed 8,192: the limit is per slab). */
int trialign_slab_buffer_bytes(int lb, int lc, int64_t *state_bytes, int64_t *flag_bytes);
int trialign_slab_score_dev(int device, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
                            int lc, int i_begin, int i_end, const void *in_state, const void *in_flags,
                            void *out_state, void *out_flags, int epoch, int score_bits, int32_t *d_out,
                            void *stream);
/* Column pipeline, the second multi-GPU decomposition of one long triple: rank g sweeps ALL planes of a contiguous range
 * of block-column pairs (columns c_begin .. c_end-1 of C; multiples of 128 except at the end); the block faces that
 * cross a range edge stream to the right neighbour's memory while the sweep runs (one face slot and one epoch-tagged
 * progress word per 64-row block, sizes from trialign_cols_buffer_bytes), so every GPU works on the same planes a few
 * block hops apart instead of waiting for whole slabs.  in_* = this rank's own buffers (NULL on the first rank),
 * out_* = the next rank's buffers as mapped with trialign_ipc_open() (NULL on the last rank).  epoch in 1..2047 must
 * differ between consecutive runs; a run must have finished on every rank before the next one starts.  Only the last
 * rank writes d_out (int32[8]).  Requires an input that is provably wrap-free at score_bits (else EINEXACT). */
int trialign_cols_buffer_bytes(int la, int lb, int64_t *face_bytes, int64_t *prog_bytes);
int trialign_cols_score_dev(int device, const uint8_t *d_a, int la, const uint8_t *d_b, int lb, const uint8_t *d_c,
                            int lc, int c_begin, int c_end, const void *in_face, const void *in_prog, void *out_face,
                            void *out_prog, int epoch, int score_bits, int32_t *d_out, void *stream);
/* Hand-over buffers must be whole cudaMalloc allocations (IPC handles name allocations, not sub-ranges). */
int trialign_device_alloc(int device, int64_t bytes, void **dptr); /* zero-filled */
int trialign_device_free(int device, void *dptr);
/* CUDA IPC plumbing for the above (one process per GPU): export a device allocation, map a peer's, unmap it. */
int trialign_ipc_export(const void *dptr, unsigned char handle[64]);
int trialign_ipc_open(int device, const unsigned char handle[64], void **dptr);
int trialign_ipc_close(void *dptr);

/* dat/{A,B,C}_seq.dat reader: one decimal digit 0-3 per line, CRLF or LF, optional trailing newline.
 * Writes at most cap symbols into buf, *n = symbols in the file (ECAPACITY if *n > cap). */
int trialign_read_seq_dat(const char *path, uint8_t *buf, int cap, int *n);
/* Writer for the same format (CRLF line ends like the reference's files when crlf != 0). */
int trialign_write_seq_dat(const char *path, const uint8_t *buf, int n, int crlf);
/* Score list in dat/Pattern1.dat's layout: ASCII signed decimals joined by '\n', no trailing newline. */
int trialign_write_scores(const char *path, const int32_t *s, int64_t n);
int trialign_read_scores(const char *path, int32_t *s, int64_t cap, int64_t *n);

/* The testbench's sequence RAMs (src/TriAlign_tb.sv:94-96,150-169): 128-bit words of 32 four-bit symbols, symbol n in
 * word n>>5 at bits [4*(n&31)+3 : 4*(n&31)]; codes A=0 T=1 C=2 G=3 N=4 (:42-46).  A word is 16 bytes here, least
 * significant byte first, so symbol n is nibble (n&1) of byte (n>>1).  pack: n symbols -> ceil(n/32) words (the rest of
 * the last word is 0 = 'A', the value the shipped testbench fills unused positions with); only bits [3:0] of a symbol
 * byte are stored.  unpack: the 4-bit codes as stored (the PE uses bits [1:0], src/PE_1cyc.v:63-66; the compute entry
 * points apply that truncation themselves).  cap_words / cap count 128-bit words / symbols; ECAPACITY if too small. */
int trialign_pack_seq_ram128(const uint8_t *sym, int n, uint8_t *ram, int cap_words);
int trialign_unpack_seq_ram128(const uint8_t *ram, int n, uint8_t *sym, int cap);
/* The same RAM images as a direct compute input (the FPGA-side format of the testbench): a sequence of
 * n symbols occupies ceil(n/32) words of 16 bytes; a batch is [n][ceil(len/32)][16] bytes per sequence array.  The batch
 * is uploaded packed (half the bytes of the one-symbol-per-byte layout) and expanded on the GPU; everything else --
 * kernels, device mask, results -- is trialign_score / trialign_score_batch. */
int trialign_score_ram128(const uint8_t *ram_a, int la, const uint8_t *ram_b, int lb, const uint8_t *ram_c, int lc,
                          int score_bits, int backend, unsigned devices, int32_t *score, int32_t final_states[7]);
int trialign_score_batch_ram128(const uint8_t *ram_a, const uint8_t *ram_b, const uint8_t *ram_c, int64_t n, int len,
                                int score_bits, int backend, unsigned devices, int32_t *scores);
/* Stimulus for the reference testbench: the `seqA_ram[w][hi:lo] <= SYM;` lines of its initial block
 * (src/TriAlign_tb.sv:423-1959), one per symbol position for A, then B, then C, each padded with `A` up to total_len
 * positions (the testbench ships total_len = 512).  A maintainer pastes / `include`s the file in place of that block
 * and sets A/B/C_LENGTH (:48-50) to replay any triple on the RTL; tools/make_tb.py does both. */
int trialign_write_tb_init(const char *path, const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c,
                           int lc, int total_len);

/* Synthetic symbols: sym = mix(seed ^ ((3*triple+which) << 32) ^ p) >> 62, splitmix64 finaliser. */
void trialign_synth(uint64_t seed, int64_t triple, int which, int len, uint8_t *out);
/* Same generator on the GPU: fills d_a/d_b/d_c [n][len] for triples first..first+n-1. */
int trialign_synth_batch_dev(int device, uint64_t seed, int64_t first, int64_t n, int len, uint8_t *d_a,
                             uint8_t *d_b, uint8_t *d_c, void *stream);

/* Bookkeeping of the last compute call made by this thread. */
int trialign_last_timing(double *seconds, double *cells); /* device time of the kernels only, cells = sum la*lb*lc */
int trialign_last_path(void);                             /* enum trialign_path */
int trialign_last_devices(void);                          /* GPUs that served the last call (0 before any call) */
int64_t trialign_last_flagged(void);                      /* triples of the last batch call that the range check sent to
                                                             the literal kernel (path FAST_THEN_LITERAL) */
int trialign_last_segments(void);                         /* segments of the last trialign_align (1 = the whole direction
                                                             cube was kept; > 1 = checkpointed recomputation) */
int64_t trialign_launch_count(void);                      /* kernels launched by this process so far */

#ifdef __cplusplus
}
#endif
#endif /* TRIALIGN_H */
