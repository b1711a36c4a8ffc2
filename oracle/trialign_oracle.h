/*
 * trialign_oracle.h -- CPU restatement of the reference's 3-sequence alignment recurrence.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product (libtrialign);
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it, and there only as the checker / the timed CPU baseline.
 *
 * Parity status: the reference (Verilog only) ships NO golden vector, expected score or
 * self-checking test for this path (TriAlign_tb.sv:337-354 only prints the score;
 * dat/Pattern1.dat is read by nothing).  The oracle is therefore pinned two ways, both
 * built from the RTL text itself:
 *   (1) this functional restatement (O1), line by line from PE_1cyc.v / TriAlign_1cyc.v;
 *   (2) oracle/rtl_model.cpp (O2), a cycle-accurate transcription of the RTL + testbench
 *       RAM port, which must agree with O1 on score, on all 7 final-cell states and on the
 *       FSM cycle-count formula.
 * "Parity unpinned by reference-held golden vectors" -- see DESIGN.md section 3.
 */
#ifndef TRIALIGN_ORACLE_H
#define TRIALIGN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* State order everywhere: M, Ix, Iy, Iz, Ixy, Iyz, Ixz  (TriAlign_1cyc.v:130 face word order). */
enum { TA_M = 0, TA_IX, TA_IY, TA_IZ, TA_IXY, TA_IYZ, TA_IXZ, TA_NSTATE };

/* Serial, literal evaluation: 49 wrapped candidates + 7 MAX7 per cell (PE_1cyc.v:164-218,139-145).
 * Symbols are bytes, only bits [1:0] are used (PE_1cyc.v:63-66).  score_bits in [2,32].
 * Returns 0, or <0 on bad arguments.  final_states / range may be NULL.
 * range[0], range[1] = min / max over every state of every interior cell (wrapped values). */
int oracle_score(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                 int score_bits, int32_t *score, int32_t final_states[7], int32_t range[2]);

/* Same function, 3D anti-diagonal wavefront with OpenMP (for long single triples). */
int oracle_score_mt(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                    int score_bits, int threads, int32_t *score, int32_t final_states[7]);

/* Same function, scheduled like the RTL's slicing (blk x blk blocks of the (j,k) plane swept over all i, faces parked
 * between blocks; TriAlign_1cyc.v:127-140): cache resident, OpenMP over the blocks of an anti-diagonal.  blk e.g. 64. */
int oracle_score_blocked(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                         int score_bits, int threads, int blk, int32_t *score, int32_t final_states[7]);

/* Uniform-length batch, structure of arrays [n][len]; OpenMP over triples. */
int oracle_score_batch(const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n, int len,
                       int score_bits, int threads, int32_t *scores);

/* Full cube dump for small cases: out[((i*(lb+1)+j)*(lc+1)+k)*7+q], i<=la, j<=lb, k<=lc. */
int oracle_cube(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc,
                int score_bits, int32_t *out);

/* Traceback of the recurrence (small cases: keeps the whole cube).  The path starts in the state
 * FINAL_MAX selects at cell (la,lb,lc) and follows, at every cell, the input the state's MAX7 instance selects (tie rule
 * of the compare-select tree, PE_1cyc.v:22-28, port wiring :139-145) until it reaches a zero border plane (i, j or k = 0).
 * moves[0..*n_moves): one byte per alignment column in forward order, bit0/1/2 = A/B/C advances (7 = M, 1 = Ix, 2 = Iy,
 * 4 = Iz, 3 = Ixy, 6 = Iyz, 5 = Ixz); start[] = the border cell the path begins at.  Returns 0, -1 bad arguments,
 * -2 out of memory, -3 internal table mismatch, -4 moves[] too small. */
int oracle_align(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                 int32_t *score, int32_t start[3], uint8_t *moves, int cap, int *n_moves);

/* The recurrence with explicit scoring parameters: scoring[5] = {MATCH, MISMATCH, GO, GE, rule}
 * with GO2 = GO<<1, GE2 = GE<<1, GOGE = GO+GE as PE_1cyc.v:57-59 derives them; rule 0 = temp_ABC as PE_1cyc.v:162 parses
 * (3*MATCH / (MATCH+MISMATCH)<<1 / 3*MISMATCH), rule 1 = sum of pairs s(a,b)+s(b,c)+s(a,c) (pic/Algorithm.png).
 * Written independently of pe_cell() above (table-driven: destination state -> predecessor offset, cost row, score),
 * so that {1,-1,2,1,0} must reproduce oracle_score() exactly -- tests/test_oracle.py checks that. */
int oracle_score_ex(const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c, int lc, int score_bits,
                    const int32_t scoring[5], int32_t *score, int32_t final_states[7]);

/* Synthetic symbol generator of SURVEY.md 8(d): splitmix64 finaliser, top 2 bits. */
void oracle_synth(uint64_t seed, int64_t triple, int which, int len, uint8_t *out);

int oracle_hw_threads(void);

#ifdef __cplusplus
}
#endif
#endif
