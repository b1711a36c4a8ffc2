#include <stdio.h>
#include "trialign.h"
int main(void) {
    uint8_t A[512], B[512], C[512]; int la = 0, lb = 0, lc = 0; int32_t score; int32_t states[7];
    if (trialign_read_seq_dat("dat/A_seq.dat", A, 512, &la) || trialign_read_seq_dat("dat/B_seq.dat", B, 512, &lb) ||
        trialign_read_seq_dat("dat/C_seq.dat", C, 512, &lc)) {
        /* no reference checkout next to the binary: a synthetic triple of the testbench's length instead */
        la = lb = lc = 64;
        trialign_synth(0x5EED0002, 0, 0, la, A); trialign_synth(0x5EED0002, 0, 1, lb, B); trialign_synth(0x5EED0002, 0, 2, lc, C);
    }
    int rc = trialign_score(A, la, B, lb, C, lc, 12, TRIALIGN_BACKEND_AUTO, 0, &score, NULL);
    if (rc) { fprintf(stderr, "%s\n", trialign_strerror(rc)); return 1; }
    printf("TriAlign Score:        \t%d\n", score);
    rc = trialign_score(A, la, B, lb, C, lc, 16, TRIALIGN_BACKEND_AUTO, 0xFF, &score, states);
    printf("%d GPUs served the call, plan %d\n", trialign_last_devices(), trialign_multi_plan(8192, 8192, 8192, 16, 8));
    trialign_scoring sc; trialign_scoring_rtl(&sc); sc.match = 2; sc.mismatch = 1; sc.triple_rule = TRIALIGN_TRIPLE_SUM_OF_PAIRS;
    rc = trialign_score_ex(A, la, B, lb, C, lc, 12, &sc, 0, &score, NULL);
    trialign_set_option("k1", "256l");
    return rc;
}
