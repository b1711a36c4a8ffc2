// host_io.cpp -- file formats and the synthetic generator (host side of the C ABI; no CUDA here).
//
// Formats follow the reference's fixtures:
//   dat/A_seq.dat, B_seq.dat, C_seq.dat : one decimal digit 0-3 per line, CRLF line ends, 64 lines
//   dat/Pattern1.dat                    : signed decimals, one per line, LF separated, no trailing newline
//   src/TriAlign_tb.sv seq?_ram         : 128-bit words of 32 four-bit symbols, and the initial block that fills them
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/trialign.h"

namespace {

int slurp(const char *path, std::string &out)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) return TRIALIGN_EIO;
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, got);
    const int bad = std::ferror(f);
    std::fclose(f);
    return bad ? TRIALIGN_EIO : TRIALIGN_OK;
}

int spill(const char *path, const std::string &s)
{
    FILE *f = std::fopen(path, "wb");
    if (!f) return TRIALIGN_EIO;
    const size_t put = std::fwrite(s.data(), 1, s.size(), f);
    const int bad = (put != s.size()) | std::fclose(f);
    return bad ? TRIALIGN_EIO : TRIALIGN_OK;
}

inline uint64_t mix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" {

int trialign_read_seq_dat(const char *path, uint8_t *buf, int cap, int *n)
{
    if (!path || !n || cap < 0 || (cap > 0 && !buf)) return TRIALIGN_EINVAL;
    std::string txt;
    const int rc = slurp(path, txt);
    if (rc) return rc;
    int count = 0;
    bool have_digit = false;  // a digit has been seen on the current line
    for (size_t p = 0; p < txt.size(); ++p) {
        const char ch = txt[p];
        if (ch >= '0' && ch <= '3') {
            if (have_digit) return TRIALIGN_EPARSE;  // two symbols on one line
            if (count < cap) buf[count] = (uint8_t)(ch - '0');
            ++count;
            have_digit = true;
        } else if (ch == '\n') {
            have_digit = false;
        } else if (ch == '\r' || ch == ' ' || ch == '\t') {
            // CRLF files (the reference's) and stray blanks are tolerated
        } else {
            return TRIALIGN_EPARSE;
        }
    }
    *n = count;
    return count > cap ? TRIALIGN_ECAPACITY : TRIALIGN_OK;
}

int trialign_write_seq_dat(const char *path, const uint8_t *buf, int n, int crlf)
{
    if (!path || n < 0 || (n > 0 && !buf)) return TRIALIGN_EINVAL;
    std::string s;
    s.reserve((size_t)n * 3);
    for (int i = 0; i < n; ++i) {
        s.push_back((char)('0' + (buf[i] & 3)));
        if (crlf) s.push_back('\r');
        s.push_back('\n');
    }
    return spill(path, s);
}

int trialign_write_scores(const char *path, const int32_t *v, int64_t n)
{
    if (!path || n < 0 || (n > 0 && !v)) return TRIALIGN_EINVAL;
    std::string s;
    s.reserve((size_t)n * 4);
    char tmp[16];
    for (int64_t i = 0; i < n; ++i) {
        if (i) s.push_back('\n');
        std::snprintf(tmp, sizeof tmp, "%d", (int)v[i]);
        s += tmp;
    }
    return spill(path, s);
}

int trialign_read_scores(const char *path, int32_t *v, int64_t cap, int64_t *n)
{
    if (!path || !n || cap < 0 || (cap > 0 && !v)) return TRIALIGN_EINVAL;
    std::string txt;
    const int rc = slurp(path, txt);
    if (rc) return rc;
    int64_t count = 0;
    const char *p = txt.c_str(), *end = p + txt.size();
    while (p < end) {
        while (p < end && (*p == '\n' || *p == '\r' || *p == ' ' || *p == '\t')) ++p;
        if (p >= end) break;
        char *q = nullptr;
        errno = 0;
        const long val = std::strtol(p, &q, 10);
        if (q == p || errno) return TRIALIGN_EPARSE;
        if (count < cap) v[count] = (int32_t)val;
        ++count;
        p = q;
    }
    *n = count;
    return count > cap ? TRIALIGN_ECAPACITY : TRIALIGN_OK;
}

int trialign_pack_seq_ram128(const uint8_t *sym, int n, uint8_t *ram, int cap_words)
{
    if (n < 0 || cap_words < 0 || (n > 0 && !sym) || (cap_words > 0 && !ram)) return TRIALIGN_EINVAL;
    const int words = (n + 31) / 32;
    if (words > cap_words) return TRIALIGN_ECAPACITY;
    std::memset(ram, 0, (size_t)words * 16);
    for (int p = 0; p < n; ++p) ram[p >> 1] |= (uint8_t)((sym[p] & 15) << (4 * (p & 1)));
    return TRIALIGN_OK;
}

int trialign_unpack_seq_ram128(const uint8_t *ram, int n, uint8_t *sym, int cap)
{
    if (n < 0 || cap < 0 || (n > 0 && !ram) || (cap > 0 && !sym)) return TRIALIGN_EINVAL;
    if (n > cap) return TRIALIGN_ECAPACITY;
    for (int p = 0; p < n; ++p) sym[p] = (uint8_t)((ram[p >> 1] >> (4 * (p & 1))) & 15);
    return TRIALIGN_OK;
}

int trialign_write_tb_init(const char *path, const uint8_t *a, int la, const uint8_t *b, int lb, const uint8_t *c,
                           int lc, int total_len)
{
    if (!path || la < 0 || lb < 0 || lc < 0 || total_len < 0 || (la > 0 && !a) || (lb > 0 && !b) || (lc > 0 && !c))
        return TRIALIGN_EINVAL;
    if (la > total_len || lb > total_len || lc > total_len) return TRIALIGN_ECAPACITY;
    static const char *const kName[5] = {"A", "T", "C", "G", "N"};  // localparams of src/TriAlign_tb.sv:42-46
    const uint8_t *seq[3] = {a, b, c};
    const int len[3] = {la, lb, lc};
    std::string s;
    s.reserve((size_t)total_len * 3 * 30);
    char line[64];
    for (int which = 0; which < 3; ++which)
        for (int p = 0; p < total_len; ++p) {
            const int code = p < len[which] ? (seq[which][p] & 15) : 0;
            if (code > 4) return TRIALIGN_EINVAL;
            const int lo = 4 * (p & 31);
            std::snprintf(line, sizeof line, "seq%c_ram[%d][%d:%d] <= %s;\n", "ABC"[which], p >> 5, lo + 3, lo, kName[code]);
            s += line;
        }
    return spill(path, s);
}

void trialign_synth(uint64_t seed, int64_t triple, int which, int len, uint8_t *out)
{
    if (!out || len <= 0) return;
    const uint64_t hi = ((uint64_t)(3 * triple + which)) << 32;
    for (int p = 0; p < len; ++p) out[p] = (uint8_t)(mix64(seed ^ hi ^ (uint64_t)p) >> 62);
}

}  // extern "C"
