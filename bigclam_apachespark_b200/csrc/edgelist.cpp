// edgelist.cpp — edge-list reader + CSR builder with the semantics the reference gets from
// GraphX: GraphLoader.edgeListFile (codes/bigclam4-7.scala:45) followed by
// collectNeighborIds(EdgeDirection.Either) (codes/bigclam4-7.scala:50).
//
//   * lines whose first non-blank character is '#', and blank lines, are skipped
//   * fields are split on runs of whitespace (tabs, spaces, CR), first two fields = src, dst
//   * a line with fewer than two fields is an error (GraphX: IllegalArgumentException("Invalid line"))
//   * every edge LINE contributes dst to src's list and src to dst's list — multiplicity is kept
//     (Email-Enron lists each pair in both directions, so every neighbour appears twice)
//   * vertex ids are arbitrary longs.  The hot path only uses ids as keys (F is keyed by id,
//     :36,:135-145), so ids are relabelled to 0..n-1 in ascending id order.
// multiplicity == 1 ("dedup") builds the simple undirected graph instead (repeated neighbours
// collapsed, self loops dropped) — what BASELINE.json means by "Email-Enron (183K edges)".
// Neighbour lists are sorted ascending (GraphX leaves the order unspecified).
#include "../../include/bigclam_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

void set_err(char *buf, int64_t len, const std::string &msg) {
    if (buf == nullptr || len <= 0) return;
    std::snprintf(buf, (size_t)len, "%s", msg.c_str());
}

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// Parses a signed decimal long starting at p (p < end); returns false if no digits.
inline bool parse_long(const char *&p, const char *end, int64_t &out) {
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) { neg = (*p == '-'); ++p; }
    if (p >= end || *p < '0' || *p > '9') return false;
    int64_t v = 0;
    while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; }
    out = neg ? -v : v;
    return true;
}

}  // namespace

extern "C" int bigclam_graph_read_edgelist(const char *path, int32_t multiplicity, bigclam_graph *out,
                                            char *errbuf, int64_t errbuf_len) {
    if (path == nullptr || out == nullptr || (multiplicity != 0 && multiplicity != 1)) {
        set_err(errbuf, errbuf_len, "bigclam_graph_read_edgelist: bad argument");
        return BIGCLAM_EINVAL;
    }
    std::memset(out, 0, sizeof(*out));
    FILE *fh = std::fopen(path, "rb");
    if (fh == nullptr) {
        set_err(errbuf, errbuf_len, std::string("cannot open ") + path);
        return BIGCLAM_EIO;
    }
    std::fseek(fh, 0, SEEK_END);
    long fsize = std::ftell(fh);
    std::fseek(fh, 0, SEEK_SET);
    std::vector<char> buf((size_t)fsize);
    if (fsize > 0 && std::fread(buf.data(), 1, (size_t)fsize, fh) != (size_t)fsize) {
        std::fclose(fh);
        set_err(errbuf, errbuf_len, std::string("short read on ") + path);
        return BIGCLAM_EIO;
    }
    std::fclose(fh);

    std::vector<int64_t> src, dst;
    const char *p = buf.data(), *end = buf.data() + buf.size();
    int64_t lineno = 0;
    while (p < end) {
        const char *eol = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        if (eol == nullptr) eol = end;
        ++lineno;
        const char *q = p;
        while (q < eol && is_space(*q)) ++q;
        if (q < eol && *q != '#') {
            int64_t a, b;
            bool ok = parse_long(q, eol, a);
            if (ok) {
                if (q < eol && !is_space(*q)) ok = false;
                while (q < eol && is_space(*q)) ++q;
                ok = ok && parse_long(q, eol, b);
                if (ok && q < eol && !is_space(*q)) ok = false;
            }
            if (!ok) {
                set_err(errbuf, errbuf_len, "Invalid line " + std::to_string(lineno) + " in " + path);
                return BIGCLAM_EIO;
            }
            src.push_back(a);
            dst.push_back(b);
        }
        p = (eol < end) ? eol + 1 : end;
    }

    const int64_t m = (int64_t)src.size();
    std::vector<int64_t> ids;
    ids.reserve((size_t)m * 2);
    ids.insert(ids.end(), src.begin(), src.end());
    ids.insert(ids.end(), dst.begin(), dst.end());
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    const int64_t n = (int64_t)ids.size();
    if (n >= (int64_t)1 << 31) {
        set_err(errbuf, errbuf_len, "more than 2^31-1 vertices");
        return BIGCLAM_EUNSUPPORTED;
    }
    auto dense = [&](int64_t id) -> int32_t {
        return (int32_t)(std::lower_bound(ids.begin(), ids.end(), id) - ids.begin());
    };

    std::vector<int64_t> rowptr((size_t)n + 1, 0);
    std::vector<int32_t> s32((size_t)m), d32((size_t)m);
    for (int64_t i = 0; i < m; ++i) {
        s32[i] = dense(src[i]);
        d32[i] = dense(dst[i]);
        ++rowptr[(size_t)s32[i] + 1];
        ++rowptr[(size_t)d32[i] + 1];
    }
    for (int64_t i = 0; i < n; ++i) rowptr[i + 1] += rowptr[i];
    std::vector<int32_t> col((size_t)rowptr[n]);
    {
        std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
        for (int64_t i = 0; i < m; ++i) {
            col[(size_t)fill[s32[i]]++] = d32[i];
            col[(size_t)fill[d32[i]]++] = s32[i];
        }
    }
    for (int64_t u = 0; u < n; ++u) std::sort(col.begin() + rowptr[u], col.begin() + rowptr[u + 1]);

    if (multiplicity == 1) {
        std::vector<int64_t> rp2((size_t)n + 1, 0);
        int64_t w = 0;
        for (int64_t u = 0; u < n; ++u) {
            int32_t prev = -1;
            for (int64_t e = rowptr[u]; e < rowptr[u + 1]; ++e) {
                const int32_t v = col[(size_t)e];
                if (v == (int32_t)u || v == prev) continue;
                col[(size_t)w++] = v;
                prev = v;
            }
            rp2[u + 1] = w;
        }
        col.resize((size_t)w);
        rowptr.swap(rp2);
    }

    out->n = n;
    out->nnz = rowptr[n];
    out->n_edge_lines = m;
    out->rowptr = (int64_t *)std::malloc(sizeof(int64_t) * ((size_t)n + 1));
    out->col = (int32_t *)std::malloc(sizeof(int32_t) * std::max<size_t>(1, col.size()));
    out->ids = (int64_t *)std::malloc(sizeof(int64_t) * std::max<size_t>(1, (size_t)n));
    if (out->rowptr == nullptr || out->col == nullptr || out->ids == nullptr) {
        bigclam_graph_free(out);
        set_err(errbuf, errbuf_len, "out of host memory");
        return BIGCLAM_ENOMEM;
    }
    std::memcpy(out->rowptr, rowptr.data(), sizeof(int64_t) * ((size_t)n + 1));
    if (!col.empty()) std::memcpy(out->col, col.data(), sizeof(int32_t) * col.size());
    if (n > 0) std::memcpy(out->ids, ids.data(), sizeof(int64_t) * (size_t)n);
    return BIGCLAM_OK;
}

extern "C" void bigclam_graph_free(bigclam_graph *g) {
    if (g == nullptr) return;
    std::free(g->rowptr);
    std::free(g->col);
    std::free(g->ids);
    std::memset(g, 0, sizeof(*g));
}
