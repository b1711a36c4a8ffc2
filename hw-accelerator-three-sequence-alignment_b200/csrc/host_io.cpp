// host_io.cpp -- file formats and the synthetic generator (host side of the C ABI; no CUDA here).
//
// Formats follow the reference's fixtures:
//   dat/A_seq.dat, B_seq.dat, C_seq.dat : one decimal digit 0-3 per line, CRLF line ends, 64 lines
//   dat/Pattern1.dat                    : signed decimals, one per line, LF separated, no trailing newline
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

void trialign_synth(uint64_t seed, int64_t triple, int which, int len, uint8_t *out)
{
    if (!out || len <= 0) return;
    const uint64_t hi = ((uint64_t)(3 * triple + which)) << 32;
    for (int p = 0; p < len; ++p) out[p] = (uint8_t)(mix64(seed ^ hi ^ (uint64_t)p) >> 62);
}

}  // extern "C"
