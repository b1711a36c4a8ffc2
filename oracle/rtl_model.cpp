/*
 * rtl_model.cpp -- O2: cycle-accurate model of the reference RTL.  TEST INFRASTRUCTURE ONLY.
 *
 * What is transcribed (file:line in /root/reference/src):
 *   PE_1cyc.v:1-32        MAX7
 *   PE_1cyc.v:55-61       scoring localparams
 *   PE_1cyc.v:148-156     outputs (M_o is Z while EN==0)
 *   PE_1cyc.v:159-218     scores + 49 candidates with the EN_i&&!EN override
 *   PE_1cyc.v:247-299     next-state of delay / symbol / EN registers
 *   PE_1cyc.v:303-430     register updates + reset values (A,B,C reset to Z)
 *   PE_2cyc.v:134-168,253-309,377-383,389-491,549-640   the 2cyc differences: registered candidates,
 *                         two-stage A/EN, delay registers gated by pp_counter==1
 *   TriAlign_1cyc.v:116-190   PE array wiring, SRAM wiring, borders, A/EN chains, FINAL_MAX
 *   TriAlign_1cyc.v:193-349   control FSM (combinational)
 *   TriAlign_1cyc.v:361-468   registers incl. the y-SRAM address/enable shift chain
 *   TriAlign_1cyc.v:472-501   sram_1024x8_t13 behavioural SRAM
 *   TriAlign_2cyc.v:119-229,232-484,496-680   the same for the 2cyc control (3-bank y-SRAM, COMPUTE/WAIT)
 *   TriAlign_tb.sv:265-271,279-333,356-421    reset, one-cycle start pulse, registered sequence-RAM read
 *
 * Modelling rules: one call to step() == one posedge clk with rst==0.  All "nxt" values are computed from
 * the pre-edge state, then committed together (every assignment in the RTL is non-blocking).  Values carry
 * an X flag; any X operand makes an arithmetic / compare / mux result X (pessimistic).
 */
#include "rtl_model.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* PE_1cyc.v:55-61 */
constexpr int MATCH = 1, MISMATCH = -1, GO = 2, GE = 1, GO2 = GO << 1, GE2 = GE << 1, GOGE = GO + GE;

struct Val {
    int32_t v;
    uint8_t x;
};
inline Val K(int32_t v) { return Val{v, 0}; }
inline Val XV() { return Val{0, 1}; }

struct Sym { /* a 4-bit symbol wire / 2-bit symbol register */
    uint8_t v;
    uint8_t x;
};

struct Word7 { /* one SRAM word / one PE output bundle: {M,Ix,Iy,Iz,Ixy,Iyz,Ixz} */
    Val s[7];
};
inline Word7 word_zero() { Word7 w; for (auto &e : w.s) e = K(0); return w; }
inline Word7 word_x() { Word7 w; for (auto &e : w.s) e = XV(); return w; }

enum { M = 0, IX, IY, IZ, IXY, IYZ, IXZ };

struct Arith {
    int sh; /* 32 - wordsize */
    inline int32_t wrap(int64_t x) const { return (int32_t)((uint32_t)x << sh) >> sh; }
    inline Val add(Val a, int c) const { return a.x ? XV() : K(wrap((int64_t)a.v + c)); }
    inline Val addv(Val a, Val c) const { return (a.x || c.x) ? XV() : K(wrap((int64_t)a.v + c.v)); }
    static inline Val mx(Val a, Val b)
    {
        if (a.x || b.x) return XV();
        return a.v > b.v ? a : b;
    }
    /* PE_1cyc.v:22-28 */
    static inline Val max7(Val g4, Val g1a, Val g1b, Val g3a, Val g2a, Val g2b, Val g3b)
    {
        Val t1 = mx(g1a, g1b), t2 = mx(g2a, g2b), t3 = mx(g3a, g3b);
        return mx(mx(t1, t2), mx(t3, g4));
    }
};

/* ---------------------------------------------------------------- PE ---------------------------------- */
struct PEIn {
    const Word7 *n1, *n2, *n3; /* [ge-1][gi-1], [ge-1][gi], [ge][gi-1] outputs */
    Sym A_i, B_i, C_i;
    uint8_t EN_i;
};

struct PE {
    /* registers */
    Val S[7];
    Val d1_1[7], d2_1[7], d1_2[7], d1_3[7];
    Val cand_d[7][7]; /* 2cyc only: *_max_*_d, [dest][source] */
    Sym A, A_d, B, C;
    uint8_t EN, EN_d;
    /* next */
    Val nS[7], nd1_1[7], nd2_1[7], nd1_2[7], nd1_3[7], ncand[7][7];
    Sym nA, nA_d, nB, nC;
    uint8_t nEN, nEN_d;

    void reset()
    {
        for (int q = 0; q < 7; ++q) {
            S[q] = d1_1[q] = d2_1[q] = d1_2[q] = d1_3[q] = K(0);
            for (int r = 0; r < 7; ++r) cand_d[q][r] = K(0);
        }
        A = A_d = B = C = Sym{0, 1}; /* 2'bz */
        EN = EN_d = 0;
    }

    inline void output(Word7 &o) const
    {
        for (int q = 0; q < 7; ++q) o.s[q] = S[q];
        if (!EN) o.s[M] = XV(); /* assign M_o = (EN==1)? M : 5'bz */
    }

    /* PE_1cyc.v:159-218: fills c[dest][source] */
    void candidates(const Arith &ar, const PEIn &in, Val c[7][7]) const
    {
        const bool sx_ab = A.x || B.x, sx_bc = B.x || C.x, sx_ac = A.x || C.x;
        const Val tAB = sx_ab ? XV() : K(A.v == B.v ? MATCH : MISMATCH);
        const Val tBC = sx_bc ? XV() : K(B.v == C.v ? MATCH : MISMATCH);
        const Val tAC = sx_ac ? XV() : K(A.v == C.v ? MATCH : MISMATCH);
        Val tABC;
        if (sx_ab) tABC = XV();
        else if (A.v != B.v) tABC = K(MISMATCH * 3);
        else if (sx_bc) tABC = XV();
        else if (B.v != C.v) tABC = K((MATCH + MISMATCH) << 1); /* `+` binds tighter than `<<` */
        else tABC = K(MATCH * 3);                                /* A==B, B==C -> A==C */
        const bool first = (in.EN_i == 1 && EN == 0);
        const Val Z0 = K(0);
        const Val *p111 = d2_1, *p110 = d1_2, *p011 = d1_1, *p101 = d1_3;
        static const int ix_cost[7] = {GO2, GE2, GOGE, GOGE, GOGE, GO2, GOGE};  /* :172-178 */
        static const int iy_cost[7] = {GO2, GOGE, GE2, GOGE, GOGE, GOGE, GO2};  /* :180-186 */
        static const int iz_cost[7] = {GO2, GOGE, GOGE, GE2, GO2, GOGE, GOGE};  /* :188-194 */
        static const int ixy_cost[7] = {GO, GE, GE, GO, GE, GO, GO};            /* :196-202 */
        static const int iyz_cost[7] = {GO, GO, GE, GE, GO, GE, GO};            /* :204-210 */
        static const int ixz_cost[7] = {GO, GE, GO, GE, GO, GO, GE};            /* :212-218 */
        for (int q = 0; q < 7; ++q) {
            c[M][q] = ar.addv(first ? Z0 : p111[q], tABC);
            c[IX][q] = ar.add(first ? Z0 : S[q], -ix_cost[q]);
            c[IY][q] = ar.add(in.n2->s[q], -iy_cost[q]);
            c[IZ][q] = ar.add(in.n3->s[q], -iz_cost[q]);
            c[IXY][q] = ar.addv(ar.add(first ? Z0 : p110[q], -ixy_cost[q]), tAB);
            c[IYZ][q] = ar.addv(ar.add(p011[q], -iyz_cost[q]), tBC);
            c[IXZ][q] = ar.addv(ar.add(first ? Z0 : p101[q], -ixz_cost[q]), tAC);
        }
    }

    static inline Val reduce(const Val c[7])
    {
        return Arith::max7(c[0], c[1], c[2], c[3], c[4], c[5], c[6]);
    }

    void eval(const Arith &ar, const PEIn &in, int variant, int pp_counter)
    {
        Val c[7][7];
        candidates(ar, in, c);
        if (variant == 1) {
            for (int q = 0; q < 7; ++q) nS[q] = reduce(c[q]);
        } else {
            for (int q = 0; q < 7; ++q) {
                nS[q] = reduce(cand_d[q]); /* MAX7 on registered candidates, PE_2cyc.v:162-168 */
                for (int r = 0; r < 7; ++r) ncand[q][r] = c[q][r];
            }
        }
        const bool adv = (variant == 1) || (pp_counter == 1); /* PE_2cyc.v:580 */
        for (int q = 0; q < 7; ++q) {
            nd1_1[q] = adv ? in.n1->s[q] : d1_1[q];
            nd2_1[q] = adv ? d1_1[q] : d2_1[q];
            nd1_2[q] = adv ? in.n2->s[q] : d1_2[q];
            nd1_3[q] = adv ? in.n3->s[q] : d1_3[q];
        }
        const Sym a2 = Sym{(uint8_t)(in.A_i.v & 3), in.A_i.x};
        if (variant == 1) {
            nA = a2;
            nEN = in.EN_i;
            nA_d = A_d;
            nEN_d = EN_d;
        } else { /* PE_2cyc.v:377-383 */
            nA = A_d;
            nA_d = a2;
            nEN = EN_d;
            nEN_d = in.EN_i;
        }
        nB = Sym{(uint8_t)(in.B_i.v & 3), in.B_i.x};
        nC = Sym{(uint8_t)(in.C_i.v & 3), in.C_i.x};
    }

    void commit(int variant)
    {
        for (int q = 0; q < 7; ++q) {
            S[q] = nS[q];
            d1_1[q] = nd1_1[q];
            d2_1[q] = nd2_1[q];
            d1_2[q] = nd1_2[q];
            d1_3[q] = nd1_3[q];
        }
        if (variant == 2) memcpy(cand_d, ncand, sizeof cand_d);
        A = nA; A_d = nA_d; B = nB; C = nC; EN = nEN; EN_d = nEN_d;
    }
};

/* ---------------------------------------------------------------- SRAM -------------------------------- */
struct Sram { /* TriAlign_1cyc.v:472-501 */
    std::vector<Word7> mem;
    Word7 Q;
    void init(int addr_bits)
    {
        mem.assign((size_t)1 << addr_bits, word_x());
        Q = word_x();
    }
    inline void clock(int CEN, int WEN, uint32_t A, const Word7 &D)
    {
        if (CEN != 0) return;
        A &= (uint32_t)(mem.size() - 1);
        if (WEN == 0) mem[A] = D;
        else Q = mem[A];
    }
};

enum { IDLE = 0, INITIAL = 1, COMPUTE = 2, OUT = 3, IN = 4, WAIT = 5 };

/* ---------------------------------------------------------------- testbench port ---------------------- */
struct TbPort { /* TriAlign_tb.sv:279-333,356-421 */
    const uint8_t *seq[3];
    int len[3];
    Sym sym[3]; /* {A,B,C}_char_ABSW: registered RAM word + registered mux index */
    int state_r, start_r;
    void reset()
    {
        for (auto &s : sym) s = Sym{0, 1};
        state_r = 0;
        start_r = 0;
    }
    inline void clock(const uint32_t addr[3])
    {
        for (int s = 0; s < 3; ++s) {
            const uint32_t ad = addr[s] & 0x7FFF;
            sym[s] = (ad < (uint32_t)len[s]) ? Sym{(uint8_t)(seq[s][ad] & 15), 0} : Sym{0, 1};
        }
        if (state_r == 0) { start_r = 1; state_r = 1; } /* S_IDLE -> S_ABSW, start pulse */
        else start_r = 0;
    }
};

/* ---------------------------------------------------------------- common array scaffolding ------------ */
struct ArrayBase {
    int P, W, variant;
    Arith ar;
    uint32_t wmask;
    std::vector<PE> pe;       /* [1..P][1..P] -> (ge-1)*P+(gi-1) */
    std::vector<Word7> out;   /* [0..P][0..P] */
    std::vector<Sym> Ao;      /* A_o[0..P][0..P] */
    std::vector<uint8_t> ENo; /* EN_o[0..P][0..P] */
    std::vector<Sym> B_i, C_i, nB_i, nC_i; /* [0..P][0..P] */
    inline int g(int ge, int gi) const { return ge * (P + 1) + gi; }
    inline PE &pe_at(int ge, int gi) { return pe[(size_t)(ge - 1) * P + (gi - 1)]; }

    void base_init(int P_, int W_, int variant_)
    {
        P = P_; W = W_; variant = variant_;
        ar.sh = 32 - W;
        wmask = (W >= 32) ? 0xFFFFFFFFu : ((1u << W) - 1u);
        pe.resize((size_t)P * P);
        for (auto &p : pe) p.reset();
        out.assign((size_t)(P + 1) * (P + 1), word_x());
        Ao.assign((size_t)(P + 1) * (P + 1), Sym{0, 1});
        ENo.assign((size_t)(P + 1) * (P + 1), 0);
        B_i.assign((size_t)(P + 1) * (P + 1), Sym{0, 0});
        C_i = B_i;
        nB_i = B_i;
        nC_i = B_i;
    }

    /* outputs of the PE registers + the A / EN chains; borders are filled by the caller first */
    void gather_outputs(Sym a_head, uint8_t EN_start)
    {
        for (int ge = 1; ge <= P; ++ge)
            for (int gi = 1; gi <= P; ++gi) {
                const PE &p = pe_at(ge, gi);
                p.output(out[g(ge, gi)]);
                Ao[g(ge, gi)] = Sym{p.A.v, p.A.x};
                ENo[g(ge, gi)] = p.EN;
            }
        ENo[g(1, 0)] = EN_start;                                  /* assign EN_o[1][0] = EN_start */
        for (int gi = 1; gi <= P; ++gi) ENo[g(0, gi)] = ENo[g(1, gi - 1)];
        Ao[g(0, 1)] = a_head;                                     /* A_o[0][1] = A_symbol / A_start */
        for (int ge = 2; ge <= P; ++ge) Ao[g(0, ge)] = Ao[g(1, ge - 1)];
    }

    void eval_pes(int pp_counter)
    {
        for (int ge = 1; ge <= P; ++ge)
            for (int gi = 1; gi <= P; ++gi) {
                PEIn in;
                in.n1 = &out[g(ge - 1, gi - 1)];
                in.n2 = &out[g(ge - 1, gi)];
                in.n3 = &out[g(ge, gi - 1)];
                in.A_i = Ao[g(ge - 1, gi)];
                in.B_i = B_i[g(ge, gi)];
                in.C_i = C_i[g(ge, gi)];
                in.EN_i = ENo[g(ge - 1, gi)];
                pe_at(ge, gi).eval(ar, in, variant, pp_counter);
            }
    }

    Val final_max() const
    { /* FINAL_MAX, TriAlign_1cyc.v:141-142 */
        const Word7 &o = out[(size_t)P * (P + 1) + P];
        return Arith::max7(o.s[IXZ], o.s[M], o.s[IX], o.s[IXY], o.s[IY], o.s[IZ], o.s[IYZ]);
    }
};

/* ---------------------------------------------------------------- TRIALIGN, 1cyc ---------------------- */
struct Trialign1 : ArrayBase {
    int A_TOTAL, ADDR_BITS, NY;
    uint32_t A_idx, B_idx, C_idx;
    uint32_t amask;
    /* registers */
    int state;
    uint32_t input_counter, compute_counter, slice_y, slice_z;
    uint8_t EN_start;
    uint32_t y_read_idx, y_write_idx;
    Val score_reg;
    uint8_t finish;
    uint32_t A_addr, B_addr, C_addr;
    std::vector<uint8_t> y_WEN, y_CEN, ny_WEN, ny_CEN;
    std::vector<uint32_t> y_A, ny_A;
    std::vector<uint8_t> z_WEN, z_CEN, nz_WEN, nz_CEN; /* [2][P+1] */
    std::vector<uint32_t> z_A, nz_A;
    std::vector<Sram> ysram, zsram; /* zsram[(ge)*(P+1)+gsz], gsz in 1..P */

    inline int zi(int b, int j) const { return b * (P + 1) + j; }

    void init(const rtl_cfg &c, int la, int lb, int lc)
    {
        base_init(c.pe_len, c.score_bits, 1);
        A_TOTAL = c.a_total_len;
        ADDR_BITS = c.sram_addr_bits;
        amask = (1u << ADDR_BITS) - 1u;
        NY = A_TOTAL + 2 * P;
        A_idx = (uint32_t)la; B_idx = (uint32_t)lb; C_idx = (uint32_t)lc;
        state = IDLE;
        input_counter = compute_counter = slice_y = slice_z = 0;
        EN_start = 0;
        y_read_idx = y_write_idx = 0;
        score_reg = K(0);
        finish = 0;
        A_addr = B_addr = C_addr = 0;
        y_WEN.assign(NY, 1); y_CEN.assign(NY, 1); y_A.assign(NY, 0);
        z_WEN.assign(2 * (P + 1), 1); z_CEN.assign(2 * (P + 1), 1); z_A.assign(2 * (P + 1), 0);
        ysram.resize(NY);
        for (auto &s : ysram) s.init(ADDR_BITS);
        zsram.resize(2 * (P + 1));
        for (auto &s : zsram) s.init(ADDR_BITS);
    }

    inline uint32_t border_00() const
    { /* wire [wordsize-1:0] border_00, :77,:149 */
        return ((y_read_idx == 0) ? (A_idx + 2 * P - 1) : (y_read_idx - 1)) & wmask;
    }

    /* one posedge clk.  Returns 1 when `finish` is set by this edge. */
    int step(TbPort &tb, Word7 *final_cell)
    {
        const uint32_t b00 = border_00();
        /* ---- borders (:155-182) ---- */
        for (int gi = 1; gi <= P; ++gi)
            out[g(0, gi)] = (slice_y == 0) ? word_zero() : ((slice_y % 2 == 1) ? zsram[zi(0, gi)].Q : zsram[zi(1, gi)].Q);
        for (int ge = 1; ge <= P; ++ge) {
            const uint32_t idx = y_read_idx + ge - 1;
            out[g(ge, 0)] = (slice_z == 0) ? word_zero() : (idx < (uint32_t)NY ? ysram[idx].Q : word_x());
        }
        out[g(0, 0)] = (slice_y == 0 || slice_z == 0) ? word_zero() : (b00 < (uint32_t)NY ? ysram[b00].Q : word_x());
        gather_outputs(tb.sym[0], EN_start);
        eval_pes(0);

        /* ---- control FSM, combinational (:193-349) ---- */
        int nstate = state;
        uint32_t n_ic = input_counter, n_cc = compute_counter, n_sy = slice_y, n_sz = slice_z;
        uint8_t n_EN_start = EN_start;
        uint32_t n_yr = y_read_idx, n_yw = y_write_idx;
        Val n_score = score_reg;
        uint8_t n_finish = finish;
        uint32_t n_Aa = A_addr, n_Ba = B_addr, n_Ca = C_addr;
        nB_i = B_i; nC_i = C_i;
        ny_WEN = y_WEN; ny_CEN = y_CEN; ny_A = y_A;
        nz_WEN = z_WEN; nz_CEN = z_CEN; nz_A = z_A;
        const uint32_t TOTAL_PE = (uint32_t)P * P;
        auto ySet = [&](std::vector<uint8_t> &v, uint32_t i, uint8_t val) { if (i < (uint32_t)NY) v[i] = val; };
        auto yASet = [&](uint32_t i, uint32_t val) { if (i < (uint32_t)NY) ny_A[i] = val & amask; };
        const uint32_t slice_y_max = ((B_idx / P) - 1) & 0xFF, slice_z_max = ((C_idx / P) - 1) & 0xFF;

        switch (state) {
        case IDLE:
            if (tb.start_r) {
                nstate = INITIAL;
                n_ic = 0;
                yASet(0, 0);
                nz_A[zi(0, 0)] = 0;
                nz_A[zi(1, 0)] = 0;
                n_Aa = n_Ba = n_Ca = 0;
            }
            break;
        case INITIAL: {
            if (input_counter >= 1) {
                const uint32_t r = (input_counter - 1) % P + 1, cidx = (input_counter - 1) / P + 1;
                if (r <= (uint32_t)P && cidx <= (uint32_t)P) {
                    nB_i[g(r, cidx)] = tb.sym[1];
                    nC_i[g(r, cidx)] = tb.sym[2];
                }
            }
            for (int j = 1; j <= P; ++j) {
                nz_WEN[zi(0, j)] = (slice_y % 2 == 1) ? 1 : 0;
                nz_WEN[zi(1, j)] = (slice_y % 2 == 1) ? 0 : 1;
                nz_CEN[zi(0, j)] = 0;
                nz_CEN[zi(1, j)] = 0;
                nz_A[zi(0, j)] = 0;
                nz_A[zi(1, j)] = 0;
            }
            for (int i = 0; i < P; ++i) {
                ySet(ny_WEN, i + y_read_idx, 1);
                ySet(ny_CEN, i + y_read_idx, 0);
                ySet(ny_CEN, i + y_write_idx, 0);
                ySet(ny_WEN, i + y_write_idx, 0);
                ySet(ny_CEN, b00, 0);
            }
            if (input_counter < TOTAL_PE) {
                for (int i = 0; i < NY; ++i) ny_A[i] = 0;
            } else {
                yASet(b00, 1);
            }
            nstate = (input_counter == TOTAL_PE) ? COMPUTE : INITIAL;
            n_ic = (input_counter == TOTAL_PE) ? 0 : ((input_counter + 1) & wmask);
            const bool b_last = (B_addr == (slice_y + 1) * P - 1);
            n_Ca = b_last ? ((C_addr == (slice_z + 1) * P - 1) ? slice_z * P : C_addr + 1) : C_addr;
            n_Ba = b_last ? slice_y * P : B_addr + 1;
            n_Aa = (input_counter == TOTAL_PE) ? 1 : 0;
            n_cc = 0;
            break;
        }
        case COMPUTE: {
            const uint32_t cc = compute_counter;
            if (cc == 0) n_EN_start = 1;
            if (cc == A_idx) n_EN_start = 0;
            if (cc < A_idx - 1) n_Aa = A_addr + 1;
            /* READ addresses (:289-304) */
            if (cc < A_idx - 2) {
                yASet(b00, cc + 2);
                yASet(y_read_idx, cc + 1);
                if (slice_y % 2 == 1) nz_A[zi(0, 1)] = (cc + 1) & amask;
                else nz_A[zi(1, 1)] = (cc + 1) & amask;
            } else if (cc >= A_idx - 2 && cc < A_idx - 1 + P) {
                yASet(y_read_idx, A_idx - 1);
                nz_A[zi(0, 1)] = (A_idx - 1) & amask;
                nz_A[zi(1, 1)] = (A_idx - 1) & amask;
            }
            /* WRITE addresses (:307-323) */
            if (cc >= (uint32_t)P && cc < P + A_idx) {
                if (slice_y % 2 == 1) nz_A[zi(1, 1)] = (cc - P) & amask;
                else nz_A[zi(0, 1)] = (cc - P) & amask;
                yASet(y_write_idx, cc - P);
            } else if (cc >= P + A_idx && cc < 2 * P + A_idx - 1) {
                nz_CEN[zi(0, 1)] = 1;
                nz_CEN[zi(1, 1)] = 1;
                ySet(ny_CEN, y_write_idx, 1);
            }
            n_cc = (cc == 2 * P + A_idx - 1) ? 0 : ((cc + 1) & wmask);
            if (cc == 2 * P + A_idx - 1) {
                n_sy = (slice_y == slice_y_max) ? 0 : ((slice_y + 1) & 0x1FF);
                n_sz = (slice_y == slice_y_max) ? ((slice_z == slice_z_max) ? 0 : ((slice_z + 1) & 0x1FF)) : slice_z;
                n_yw = (y_write_idx == A_idx + P) ? 0 : ((y_write_idx + P) & 0x3FF);
                if (slice_z > 0) n_yr = (y_read_idx >= A_idx + P) ? 0 : ((y_read_idx + P) & 0x3FF);
                else n_yr = y_read_idx;
                nstate = (slice_y == slice_y_max && slice_z == slice_z_max) ? OUT : INITIAL;
                n_Aa = 0;
                n_Ba = (slice_y == slice_y_max) ? 0 : (slice_y + 1) * P;
                n_Ca = (slice_y == slice_y_max) ? ((slice_z == slice_z_max) ? 0 : (slice_z + 1) * P) : C_addr;
                if (slice_z == slice_z_max && slice_y == slice_y_max) {
                    n_score = final_max();
                    n_finish = 1;
                    if (final_cell) *final_cell = out[g(P, P)];
                }
            }
            break;
        }
        default: break; /* OUT has no arm in the 1cyc case statement: everything holds */
        }

        /* ---- SRAMs see pre-edge CEN/WEN/A/D (:129-140, :493-500) ---- */
        for (int b = 0; b <= 1; ++b)
            for (int gsz = 1; gsz <= P; ++gsz)
                zsram[zi(b, gsz)].clock(z_CEN[zi(b, gsz)], z_WEN[zi(b, gsz)], z_A[zi(b, gsz)], out[g(P, gsz)]);
        for (int gsy = 0; gsy < NY; ++gsy) {
            if (y_CEN[gsy]) continue;
            ysram[gsy].clock(0, y_WEN[gsy], y_A[gsy], out[g(gsy % P + 1, P)]); /* y_D_i[gi*P+ge] = row ge+1 */
        }

        /* ---- SRAM control registers (:361-423) ---- */
        {
            std::vector<uint8_t> &cy_CEN = y_CEN;
            std::vector<uint32_t> &cy_A = y_A;
            /* iterate downwards so y_*[i-1] is still the pre-edge value when read */
            for (int i = NY - 1; i >= 0; --i) {
                const uint32_t ui = (uint32_t)i;
                if (ui > y_write_idx && ui < y_write_idx + P && i >= 1) {
                    cy_A[i] = cy_A[i - 1];
                    cy_CEN[i] = cy_CEN[i - 1];
                } else if (ui == y_read_idx || ui == y_write_idx || ui == b00) {
                    cy_A[i] = ny_A[i];
                    cy_CEN[i] = ny_CEN[i];
                } else if (ui > y_read_idx && ui < y_read_idx + P && i >= 1) {
                    cy_A[i] = cy_A[i - 1];
                    cy_CEN[i] = ny_CEN[i];
                } else {
                    cy_A[i] = ny_A[i];
                    cy_CEN[i] = ny_CEN[i];
                }
            }
            y_WEN = ny_WEN;
            for (int b = 0; b <= 1; ++b) {
                for (int j = P; j >= 2; --j) {
                    z_A[zi(b, j)] = z_A[zi(b, j - 1)];
                    z_CEN[zi(b, j)] = z_CEN[zi(b, j - 1)];
                }
                z_A[zi(b, 1)] = nz_A[zi(b, 1)];
                z_CEN[zi(b, 1)] = nz_CEN[zi(b, 1)];
                /* index 0 is never shifted into nor connected to an SRAM; keep nxt for completeness */
            }
            z_WEN = nz_WEN;
        }

        /* ---- testbench registers use the pre-edge addresses ---- */
        const uint32_t addr[3] = {A_addr, B_addr, C_addr};
        tb.clock(addr);

        /* ---- commit ---- */
        for (auto &p : pe) p.commit(1);
        state = nstate;
        input_counter = n_ic; compute_counter = n_cc; slice_y = n_sy; slice_z = n_sz;
        B_i.swap(nB_i); C_i.swap(nC_i);
        EN_start = n_EN_start;
        y_read_idx = n_yr; y_write_idx = n_yw;
        score_reg = n_score;
        finish = n_finish;
        A_addr = n_Aa & 0x7FFF; B_addr = n_Ba & 0x7FFF; C_addr = n_Ca & 0x7FFF;
        return finish;
    }
};

/* ---------------------------------------------------------------- TRIALIGN, 2cyc ---------------------- */
struct Trialign2 : ArrayBase {
    int A_TOTAL, ADDR_BITS;
    uint32_t A_idx, B_idx, C_idx;
    uint32_t amask_z, amask_y;
    /* registers */
    int state;
    uint32_t input_counter, compute_counter, slice_y, slice_z, pp_counter;
    uint8_t EN_start;
    Sym A_start;
    uint32_t y_read_idx, y_write_idx;
    Val score_reg;
    uint8_t finish;
    uint32_t A_addr, B_addr, C_addr;
    std::vector<uint8_t> y_WEN, y_CEN, ny_WEN, ny_CEN; /* [3][2][P+1] */
    std::vector<uint32_t> y_A, ny_A;
    std::vector<uint8_t> z_WEN, z_CEN, nz_WEN, nz_CEN; /* [2][P+1] */
    std::vector<uint32_t> z_A, nz_A;
    std::vector<Sram> ysram, zsram;

    inline int zi(int b, int j) const { return b * (P + 1) + j; }
    inline int yi(int i, int j, int k) const { return (i * 2 + j) * (P + 1) + k; }

    void init(const rtl_cfg &c, int la, int lb, int lc)
    {
        base_init(c.pe_len, c.score_bits, 2);
        A_TOTAL = c.a_total_len;
        ADDR_BITS = c.sram_addr_bits;
        amask_z = (1u << ADDR_BITS) - 1u;
        amask_y = (1u << (ADDR_BITS + 4)) - 1u;
        A_idx = (uint32_t)la; B_idx = (uint32_t)lb; C_idx = (uint32_t)lc;
        state = IDLE;
        input_counter = compute_counter = slice_y = slice_z = pp_counter = 0;
        EN_start = 0;
        A_start = Sym{0, 1}; /* A_start <= 1'bz (TriAlign_2cyc.v:648) */
        y_read_idx = 16;     /* :650 */
        y_write_idx = 0;
        score_reg = K(0);
        finish = 0;
        A_addr = B_addr = C_addr = 0;
        const int ny = 3 * 2 * (P + 1);
        y_WEN.assign(ny, 1); y_CEN.assign(ny, 1); y_A.assign(ny, 0);
        z_WEN.assign(2 * (P + 1), 1); z_CEN.assign(2 * (P + 1), 1); z_A.assign(2 * (P + 1), 0);
        ysram.resize(ny);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k <= P; ++k) ysram[yi(i, j, k)].init(i < 2 ? ADDR_BITS + 4 : ADDR_BITS); /* :141-157 */
        zsram.resize(2 * (P + 1));
        for (auto &s : zsram) s.init(ADDR_BITS);
    }

    inline uint32_t border_00() const
    {
        return ((y_read_idx == 0) ? (A_idx + 2 * P - 1) : (y_read_idx - 1)) & wmask;
    }
    /* :178-180: "group 2 discrimination" uses fixed bit positions 6..9 of the index against B_idx */
    inline bool bank2(uint32_t idx) const
    {
        for (int b = 6; b <= 9; ++b)
            if (((idx >> b) & 1u) && ((B_idx >> b) & 1u)) return true;
        return false;
    }
    static inline int bit(uint32_t v, int b) { return (int)((v >> b) & 1u); }

    int step(TbPort &tb, Word7 *final_cell)
    {
        const uint32_t b00 = border_00();
        const bool b00_b = bank2(b00), yr_b = bank2(y_read_idx), yw_b = bank2(y_write_idx);
        const int yr4 = bit(y_read_idx, 4), yr3 = bit(y_read_idx, 3);
        const int yw4 = bit(y_write_idx, 4), yw3 = bit(y_write_idx, 3);
        const int b4 = bit(b00, 4), b3 = bit(b00, 3);
        const int r_bank = yr_b ? 2 : yr4, w_bank = yw_b ? 2 : yw4, c_bank = b00_b ? 2 : b4;
        const uint32_t yr_hi = ((y_read_idx >> 5) & 31u) << 9, yw_hi = ((y_write_idx >> 5) & 31u) << 9,
                       b00_hi = ((b00 >> 5) & 31u) << 9;

        /* ---- borders (:186-221) ---- */
        for (int gi = 1; gi <= P; ++gi)
            out[g(0, gi)] = (slice_y == 0) ? word_zero() : ((slice_y % 2 == 1) ? zsram[zi(0, gi)].Q : zsram[zi(1, gi)].Q);
        for (int ge = 1; ge <= P; ++ge)
            out[g(ge, 0)] = (slice_z == 0) ? word_zero() : ysram[yi(r_bank, yr3, ge)].Q;
        out[g(0, 0)] = (slice_y == 0 || slice_z == 0) ? word_zero() : ysram[yi(c_bank, b3, P)].Q;
        gather_outputs(A_start, EN_start);
        eval_pes((int)pp_counter);

        /* ---- control FSM, combinational (:232-484) ---- */
        int nstate = state;
        uint32_t n_ic = input_counter, n_cc = compute_counter, n_sy = slice_y, n_sz = slice_z, n_pp = pp_counter;
        uint8_t n_EN_start = EN_start;
        Sym n_A_start = A_start;
        uint32_t n_yr = y_read_idx, n_yw = y_write_idx;
        Val n_score = score_reg;
        uint8_t n_finish = finish;
        uint32_t n_Aa = A_addr, n_Ba = B_addr, n_Ca = C_addr;
        nB_i = B_i; nC_i = C_i;
        ny_WEN = y_WEN; ny_CEN = y_CEN; ny_A = y_A;
        nz_WEN = z_WEN; nz_CEN = z_CEN; nz_A = z_A;
        const uint32_t TOTAL_PE = (uint32_t)P * P;
        const uint32_t slice_y_max = ((B_idx / P) - 1) & 0xFF, slice_z_max = ((C_idx / P) - 1) & 0xFF;

        switch (state) {
        case IDLE:
            if (tb.start_r) {
                nstate = INITIAL;
                n_ic = 0;
                nz_A[zi(0, 0)] = 0;
                nz_A[zi(1, 0)] = 0;
                n_Aa = n_Ba = n_Ca = 0;
            }
            break;
        case INITIAL: {
            if (input_counter >= 1) {
                const uint32_t r = (input_counter - 1) % P + 1, cidx = (input_counter - 1) / P + 1;
                if (r <= (uint32_t)P && cidx <= (uint32_t)P) {
                    nB_i[g(r, cidx)] = tb.sym[1];
                    nC_i[g(r, cidx)] = tb.sym[2];
                }
            }
            for (int j = 1; j <= P; ++j) {
                nz_WEN[zi(0, j)] = (slice_y % 2 == 1) ? 1 : 0;
                nz_WEN[zi(1, j)] = (slice_y % 2 == 1) ? 0 : 1;
                nz_CEN[zi(0, j)] = 0;
                nz_CEN[zi(1, j)] = 0;
                nz_A[zi(0, j)] = 0;
                nz_A[zi(1, j)] = 0;
            }
            for (int k = 1; k <= P; ++k) { /* :305-322, read first then write (write wins on a clash) */
                ny_WEN[yi(r_bank, yr3, k)] = 1;
                ny_CEN[yi(r_bank, yr3, k)] = 0;
                ny_WEN[yi(w_bank, yw3, k)] = 0;
                ny_CEN[yi(w_bank, yw3, k)] = 0;
            }
            ny_WEN[yi(c_bank, b3, P)] = 1; /* :325-331 */
            ny_CEN[yi(c_bank, b3, P)] = 0;
            if (input_counter < TOTAL_PE - 1) { /* :334-352 */
                for (int k = 0; k <= P; ++k) ny_A[yi(r_bank, yr3, k)] = yr_hi & amask_y;
            } else {
                ny_A[yi(c_bank, b3, P)] = b00_hi & amask_y;
            }
            nstate = (input_counter == TOTAL_PE) ? COMPUTE : INITIAL;
            n_ic = (input_counter == TOTAL_PE) ? 0 : ((input_counter + 1) & wmask);
            const bool b_last = (B_addr == (slice_y + 1) * P - 1);
            n_Ca = b_last ? ((C_addr == (slice_z + 1) * P - 1) ? slice_z * P : C_addr + 1) : C_addr;
            n_Ba = b_last ? slice_y * P : B_addr + 1;
            n_Aa = (input_counter >= TOTAL_PE - 1) ? 1 : 0;
            n_cc = 0;
            n_A_start = tb.sym[0];
            n_pp = (pp_counter == 1) ? 0 : ((pp_counter + 1) & 7);
            break;
        }
        case COMPUTE: {
            const uint32_t cc = compute_counter;
            if (cc < A_idx - 1) { /* READ addresses :368-390 */
                ny_A[yi(r_bank, yr3, 1)] = (cc + yr_hi) & amask_y;
                ny_A[yi(c_bank, b3, P)] = (cc + 1 + b00_hi) & amask_y;
                if (slice_y % 2 == 1) nz_A[zi(0, 1)] = cc & amask_z;
                else nz_A[zi(1, 1)] = cc & amask_z;
            } else if (cc >= A_idx - 1 && cc < A_idx + P) {
                ny_A[yi(r_bank, yr3, 1)] = (A_idx - 1 + yr_hi) & amask_y;
                nz_A[zi(0, 1)] = (A_idx - 1) & amask_z;
                nz_A[zi(1, 1)] = (A_idx - 1) & amask_z;
            }
            if (cc >= (uint32_t)P && cc < P + A_idx) { /* WRITE addresses :406-431 */
                if (slice_y % 2 == 1) nz_A[zi(1, 1)] = (cc - P) & amask_z;
                else nz_A[zi(0, 1)] = (cc - P) & amask_z;
                ny_A[yi(w_bank, yw3, 1)] = (cc - P + yw_hi) & amask_y;
            } else if (cc >= P + A_idx && cc < 2 * P + A_idx - 1) {
                nz_CEN[zi(0, 1)] = 1;
                nz_CEN[zi(1, 1)] = 1;
                ny_CEN[yi(w_bank, yw3, 1)] = 1;
            }
            nstate = WAIT;
            n_pp = 1;
            break;
        }
        case OUT: nstate = OUT; break;
        case WAIT: {
            const uint32_t cc = compute_counter;
            n_cc = (cc == 2 * P + A_idx - 1) ? 0 : ((cc + 1) & wmask);
            n_pp = 0;
            if (cc == 2 * P + A_idx - 1) {
                n_sy = (slice_y == slice_y_max) ? 0 : ((slice_y + 1) & 0x1FF);
                n_sz = (slice_y == slice_y_max) ? ((slice_z == slice_z_max) ? 0 : ((slice_z + 1) & 0x1FF)) : slice_z;
                n_yw = (y_write_idx == A_idx + P) ? 0 : ((y_write_idx + P) & 0x3FF);
                n_yr = (y_read_idx >= A_idx + P) ? 0 : ((y_read_idx + P) & 0x3FF);
                nstate = (slice_y == slice_y_max && slice_z == slice_z_max) ? OUT : INITIAL;
                n_Aa = 0;
                n_Ba = (slice_y == slice_y_max) ? 0 : (slice_y + 1) * P;
                n_Ca = (slice_y == slice_y_max) ? ((slice_z == slice_z_max) ? 0 : (slice_z + 1) * P) : C_addr;
                if (slice_z == slice_z_max && slice_y == slice_y_max) {
                    n_score = final_max();
                    n_finish = 1;
                    if (final_cell) *final_cell = out[g(P, P)];
                }
            } else {
                if (cc < A_idx - 1) {
                    n_A_start = tb.sym[0];
                    n_Aa = A_addr + 1;
                }
                nstate = COMPUTE;
                if (cc == 0) n_EN_start = 1;
                if (cc == A_idx) n_EN_start = 0;
            }
            break;
        }
        default: break;
        }

        /* ---- SRAMs see pre-edge control and data ---- */
        for (int b = 0; b <= 1; ++b)
            for (int gsz = 1; gsz <= P; ++gsz)
                zsram[zi(b, gsz)].clock(z_CEN[zi(b, gsz)], z_WEN[zi(b, gsz)], z_A[zi(b, gsz)], out[g(P, gsz)]);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 1; k <= P; ++k) {
                    const int id = yi(i, j, k);
                    if (y_CEN[id]) continue;
                    const uint32_t addr = (i == 2) ? (y_A[id] & 0x1FFu) : y_A[id]; /* :155: y_A_i[2][..][8:0] */
                    ysram[id].clock(0, y_WEN[id], addr, out[g(k, P)]);              /* :162: row gsy, last column */
                }

        /* ---- SRAM control registers (:496-631) ---- */
        if (state == COMPUTE) {
            for (int i = 0; i <= 1; ++i)
                for (int j = 0; j <= 1; ++j)
                    for (int k = 0; k <= P; ++k) y_WEN[yi(i, j, k)] = ny_WEN[yi(i, j, k)]; /* bank 2 holds (:524-529) */
            z_WEN = nz_WEN;
            for (int i = 0; i <= 2; ++i)
                for (int j = 0; j <= 1; ++j)
                    for (int k = P; k >= 0; --k) { /* downwards: [k-1] is still the pre-edge value */
                        const int id = yi(i, j, k);
                        const bool wsel = yw_b ? (i == 2 && yw3 == j) : (yw4 == i && yw3 == j);
                        const bool rsel = yr_b ? (i == 2 && yr3 == j) : (yr4 == i && yr3 == j);
                        if (k > 1 && wsel) y_CEN[id] = y_CEN[id - 1];
                        else y_CEN[id] = ny_CEN[id];
                        if (k > 1 && (wsel || rsel)) y_A[id] = y_A[id - 1];
                        else y_A[id] = ny_A[id];
                    }
            for (int b = 0; b <= 1; ++b) {
                for (int j = P; j >= 2; --j) {
                    z_A[zi(b, j)] = z_A[zi(b, j - 1)];
                    z_CEN[zi(b, j)] = z_CEN[zi(b, j - 1)];
                }
                z_A[zi(b, 1)] = nz_A[zi(b, 1)];
                z_CEN[zi(b, 1)] = nz_CEN[zi(b, 1)];
            }
        } else {
            z_WEN = nz_WEN; z_CEN = nz_CEN; z_A = nz_A;
            y_WEN = ny_WEN; y_CEN = ny_CEN; y_A = ny_A;
        }

        const uint32_t addr[3] = {A_addr, B_addr, C_addr};
        tb.clock(addr);

        for (auto &p : pe) p.commit(2);
        state = nstate;
        input_counter = n_ic; compute_counter = n_cc; slice_y = n_sy; slice_z = n_sz; pp_counter = n_pp;
        B_i.swap(nB_i); C_i.swap(nC_i);
        EN_start = n_EN_start;
        A_start = n_A_start;
        y_read_idx = n_yr; y_write_idx = n_yw;
        score_reg = n_score;
        finish = n_finish;
        A_addr = n_Aa & 0x7FFF; B_addr = n_Ba & 0x7FFF; C_addr = n_Ca & 0x7FFF;
        return finish;
    }
};

int ceil_log2(int v)
{
    int b = 0;
    while ((1 << b) < v) ++b;
    return b;
}

/* SURVEY.md 8(a) "Validity domain" */
int in_domain(const rtl_cfg &c, int la, int lb, int lc)
{
    const int P = c.pe_len, W = c.score_bits;
    if (la < 2 * 1 || lb < P || lc < P) return 0;
    if (lb % P || lc % P) return 0;                    /* remainder silently dropped (:50-51) */
    if (W < 7) return 0;                               /* input_counter must reach P*P = 64 at P = 8 */
    if ((int64_t)P * P >= ((int64_t)1 << W)) return 0;
    if ((int64_t)la + 2 * P - 1 >= ((int64_t)1 << W)) return 0; /* compute_counter is W bits */
    if (la > (1 << c.sram_addr_bits)) return 0;        /* face depth */
    if (la > c.a_total_len) return 0;
    if (la + P > 1023) return 0;                       /* y_read_idx / y_write_idx are 10 bits */
    if (lb / P - 1 > 255 || lc / P - 1 > 255) return 0; /* slice_*_max_idx are 8 bits */
    if (c.variant == 1) {
        if (lc > P && la < lb) return 0;               /* y ring has A_idx/P+2 groups */
    } else {
        /* TriAlign_2cyc selects its y-SRAM bank from fixed bit positions of the ring indices
         * (TriAlign_2cyc.v:178-180,205-219, y_read_idx resets to 16 at :650): built for A=B=C in {64,128,256,512};
         * e.g. 24 and 32 put the read and the write group into the same single-port SRAM and yield X. */
        if (la != lb || lb != lc) return 0;
        if (la % 64) return 0;
    }
    return 1;
}

} // namespace

extern "C" {

void rtl_default_cfg(rtl_cfg *cfg, int variant)
{
    cfg->variant = variant;
    cfg->pe_len = 8;
    cfg->score_bits = 12;
    cfg->a_total_len = 0;
    cfg->sram_addr_bits = 0;
    cfg->strict = 1;
    cfg->max_cycles = 0;
}

int64_t rtl_cycle_formula(int variant, int P, int la, int lb, int lc)
{
    const int64_t slices = (int64_t)(lb / P) * (lc / P);
    const int64_t per = (variant == 1) ? ((int64_t)P * P + 1 + la + 2 * P) : ((int64_t)P * P + 1 + 2 * ((int64_t)la + 2 * P));
    return slices * per + 2;
}

int rtl_model_run(const rtl_cfg *cfg_in, const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c,
                  int lc, int32_t *score, int32_t final_states[7], int64_t *cycles)
{
    if (!cfg_in || !a || !b || !c || !score || la <= 0 || lb <= 0 || lc <= 0) return RTL_EARG;
    rtl_cfg cfg = *cfg_in;
    if (cfg.variant != 1 && cfg.variant != 2) return RTL_EARG;
    if (cfg.pe_len < 1 || cfg.pe_len > 64 || cfg.score_bits < 2 || cfg.score_bits > 30) return RTL_EARG;
    if (cfg.a_total_len <= 0) cfg.a_total_len = std::max(512, la);
    if (cfg.sram_addr_bits <= 0) cfg.sram_addr_bits = std::max(9, ceil_log2(la));
    if (cfg.sram_addr_bits > 20) return RTL_EARG;
    if (cfg.strict && !in_domain(cfg, la, lb, lc)) return RTL_EDOMAIN;
    const int64_t limit = cfg.max_cycles > 0 ? cfg.max_cycles : 3000000;

    TbPort tb;
    tb.seq[0] = a; tb.seq[1] = b; tb.seq[2] = c;
    tb.len[0] = la; tb.len[1] = lb; tb.len[2] = lc;
    tb.reset();

    Word7 fin = word_x();
    Val sc = XV();
    int64_t n = 0;
    bool done = false;
    if (cfg.variant == 1) {
        auto dut = std::make_unique<Trialign1>();
        dut->init(cfg, la, lb, lc);
        while (n < limit) {
            ++n;
            if (dut->step(tb, &fin)) { done = true; break; }
        }
        sc = dut->score_reg;
    } else {
        auto dut = std::make_unique<Trialign2>();
        dut->init(cfg, la, lb, lc);
        while (n < limit) {
            ++n;
            if (dut->step(tb, &fin)) { done = true; break; }
        }
        sc = dut->score_reg;
    }
    if (cycles) *cycles = n;
    if (!done) return RTL_ETIMEOUT;
    if (final_states)
        for (int q = 0; q < 7; ++q) final_states[q] = fin.s[q].x ? INT32_MIN : fin.s[q].v;
    if (sc.x) return RTL_EX;
    *score = sc.v;
    return RTL_OK;
}

int rtl_model_run_batch(const rtl_cfg *cfg, const uint8_t *a, const uint8_t *b, const uint8_t *c, int64_t n,
                        int len, int threads, int32_t *scores, int64_t *total_cycles)
{
    if (!cfg || n < 0 || len <= 0 || !scores) return RTL_EARG;
    int rc = 0;
    int64_t tot = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(+ : tot)
#endif
    for (int64_t t = 0; t < n; ++t) {
        int64_t cyc = 0;
        int r = rtl_model_run(cfg, a + t * len, len, b + t * len, len, c + t * len, len, &scores[t], nullptr, &cyc);
        tot += cyc;
        if (r) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            rc = r;
        }
    }
    if (total_cycles) *total_cycles = tot;
    return rc;
}

} // extern "C"
