/*
 * rtl_model.h -- O2, a hand-written cycle-accurate C++ model of the reference RTL.
 *
 * TEST INFRASTRUCTURE ONLY (checker + the "RTL simulation" CPU baseline of bench.py).
 * No Verilog simulator exists in this image (iverilog / verilator / vcs ... all absent), so the
 * RTL is transcribed register by register; every clocked register updates from pre-edge values,
 * X/Z are modelled pessimistically at word granularity (an X operand makes the result X), and the
 * run fails loudly if the final score is X.
 *
 * Variant 1: src/PE_1cyc.v + src/TriAlign_1cyc.v   (1 clock per wavefront step)
 * Variant 2: src/PE_2cyc.v + src/TriAlign_2cyc.v   (2 clocks per wavefront step)
 * Environment: the sequence-RAM port of src/TriAlign_tb.sv:279-421 (1-cycle registered read,
 * one-cycle start pulse two clocks after reset).
 */
#ifndef TRIALIGN_RTL_MODEL_H
#define TRIALIGN_RTL_MODEL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtl_cfg {
    int variant;        /* 1 = *_1cyc, 2 = *_2cyc */
    int pe_len;         /* PE_LEN, tb: 8 (TriAlign_tb.sv:54) */
    int score_bits;     /* SCORE_BITS, tb: 12 (:56) */
    int a_total_len;    /* A_TOTAL_LEN, tb: 512 (:51); 0 = max(512, la) */
    int sram_addr_bits; /* SRAM_ADDR_BITS, tb: 9 (:55); 0 = max(9, ceil(log2(la))) */
    int strict;         /* 1: refuse inputs outside the RTL's validity domain */
    int64_t max_cycles; /* 0 = 30 ms of simulated time at 100 MHz (TriAlign_tb.sv:13,273-276) */
} rtl_cfg;

enum {
    RTL_OK = 0,
    RTL_EARG = -1,      /* bad pointer / parameter */
    RTL_EDOMAIN = -2,   /* strict mode: outside the RTL's validity domain */
    RTL_ETIMEOUT = -3,  /* finish never rose */
    RTL_EX = -4         /* Score is X (uninitialised / contended storage reached the result) */
};

void rtl_default_cfg(rtl_cfg *cfg, int variant);

/* Runs reset + one alignment.  symbols: bytes, 4-bit codes as the tb RAM holds them (A=0,T=1,C=2,G=3,N=4);
 * the PE keeps bits [1:0].  cycles = posedges from the first non-reset edge through the edge that sets
 * `finish`.  final_states = registered outputs of PE[P][P] at that edge (M,Ix,Iy,Iz,Ixy,Iyz,Ixz). */
int rtl_model_run(const rtl_cfg *cfg, const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c,
                  int lc, int32_t *score, int32_t final_states[7], int64_t *cycles);

/* Closed-form cycle count of the FSM, for cross-checking. */
int64_t rtl_cycle_formula(int variant, int pe_len, int la, int lb, int lc);

/* One triple per thread (OpenMP); returns the first non-zero rc.  total_cycles may be NULL. */
int rtl_model_run_batch(const rtl_cfg *cfg, const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n,
                        int len, int threads, int32_t *scores, int64_t *total_cycles);

#ifdef __cplusplus
}
#endif
#endif
