// sx_host.cpp -- host-side pieces of the drop-in surface that need no GPU:
// the Matrix Market -> CSR loader and the row-block partitioner.
//
// The loader reproduces the OBSERVABLE behaviour of the reference's
// read_suitsparse_matrix(..., CSC) + CSC_2_CSR pair (src/sparse_helper.h:112-259,
// 475-509; banner/size rules of src/mmio.h:254-367) with a different mechanism: the
// file is read once into memory and tokenised by hand (the reference calls fscanf per
// entry), and the (column,row) qsort + CSC->CSR sweep is replaced by two stable
// counting sorts, which give the same CSR: rows in order, columns ascending within a
// row.  Entries with equal (row,col) keep file order here; the reference's qsort
// leaves their order unspecified.
//
// Deliberate differences, all on malformed input where the reference has undefined
// or silent behaviour: truncated files, unparsable tokens and indices beyond the
// declared size return an error instead of reusing stale values / writing out of
// bounds.
#include "../../include/sextans_b200.h"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace sxhost {

thread_local std::string g_err;

static int fail(int status, const std::string &msg) {
    g_err = msg;
    return status;
}

struct Banner {
    bool coordinate = false, pattern = false, complex_ = false, symmetric = false;
};

static std::string lower(std::string s) {
    for (auto &ch : s) ch = (char)std::tolower((unsigned char)ch);
    return s;
}

class Cursor {
  public:
    Cursor(const char *b, const char *e) : p_(b), e_(e) {}
    bool eof() const { return p_ >= e_; }
    // one text line without its terminator; false at end of data
    bool line(std::string *out) {
        if (p_ >= e_) return false;
        const char *nl = (const char *)memchr(p_, '\n', (size_t)(e_ - p_));
        const char *stop = nl ? nl : e_;
        out->assign(p_, stop);
        p_ = nl ? nl + 1 : e_;
        return true;
    }
    void skip_ws() {
        while (p_ < e_ && std::isspace((unsigned char)*p_)) ++p_;
    }
    bool integer(long *v) {
        skip_ws();
        if (p_ >= e_) return false;
        const char *q = p_;
        bool neg = false;
        if (*q == '+' || *q == '-') { neg = *q == '-'; ++q; }
        if (q >= e_ || !std::isdigit((unsigned char)*q)) return false;
        long acc = 0;
        while (q < e_ && std::isdigit((unsigned char)*q)) { acc = acc * 10 + (*q - '0'); ++q; }
        *v = neg ? -acc : acc;
        p_ = q;
        return true;
    }
    // value token converted exactly as scanf's %f / %lg would (strtof / strtod)
    template <typename T> bool real(T *v) {
        skip_ws();
        if (p_ >= e_) return false;
        char *endp = nullptr;
        if (sizeof(T) == 4) *v = (T)std::strtof(p_, &endp);
        else *v = (T)std::strtod(p_, &endp);
        if (endp == p_) return false;
        p_ = endp;
        return true;
    }

  private:
    const char *p_, *e_;
};

static int parse_banner(Cursor &cur, Banner *b) {
    std::string first;
    if (!cur.line(&first)) return fail(SX_ERR_IO, "empty file");
    char tok[5][64];
    if (std::sscanf(first.c_str(), "%63s %63s %63s %63s %63s", tok[0], tok[1], tok[2], tok[3], tok[4]) != 5)
        return fail(SX_ERR_FORMAT, "Matrix Market banner needs five fields");
    if (std::strncmp(tok[0], "%%MatrixMarket", 14) != 0) return fail(SX_ERR_FORMAT, "missing %%MatrixMarket banner");
    const std::string object = lower(tok[1]), format = lower(tok[2]), field = lower(tok[3]), sym = lower(tok[4]);
    if (object != "matrix") return fail(SX_ERR_FORMAT, "banner object is not 'matrix'");
    if (format == "coordinate") b->coordinate = true;
    else if (format != "array") return fail(SX_ERR_FORMAT, "unknown banner format '" + format + "'");
    if (field == "pattern") b->pattern = true;
    else if (field == "complex") b->complex_ = true;
    else if (field != "real" && field != "integer") return fail(SX_ERR_FORMAT, "unknown banner field '" + field + "'");
    if (sym == "symmetric") b->symmetric = true;  // hermitian / skew-symmetric are read as stored
    else if (sym != "general" && sym != "hermitian" && sym != "skew-symmetric")
        return fail(SX_ERR_FORMAT, "unknown banner symmetry '" + sym + "'");
    return SX_OK;
}

static int parse_size(Cursor &cur, long *M, long *K, long *nz) {
    std::string ln;
    do {
        if (!cur.line(&ln)) return fail(SX_ERR_IO, "no size line");
    } while (!ln.empty() && ln[0] == '%');
    if (std::sscanf(ln.c_str(), "%ld %ld %ld", M, K, nz) == 3) return SX_OK;
    // blank line(s) before the size line: keep reading whitespace-separated integers
    if (!cur.integer(M) || !cur.integer(K) || !cur.integer(nz)) return fail(SX_ERR_IO, "bad size line");
    return SX_OK;
}

template <typename T> static bool is_plus_zero(T v) {
    unsigned char z[sizeof(T)] = {0};
    return std::memcmp(&v, z, sizeof(T)) == 0;
}

template <typename T>
static int load(const char *path, int *M_out, int *K_out, int64_t *nnz_out, int32_t **rowptr_out,
                int32_t **colidx_out, T **val_out) {
    if (!path || !M_out || !K_out || !nnz_out || !rowptr_out || !colidx_out || !val_out)
        return fail(SX_ERR_INVALID, "null argument");
    *rowptr_out = nullptr; *colidx_out = nullptr; *val_out = nullptr;
    FILE *f = std::fopen(path, "rb");
    if (!f) return fail(SX_ERR_IO, std::string("Could not open ") + path);
    std::vector<char> text;
    {
        char buf[1 << 16];
        size_t got;
        while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.insert(text.end(), buf, buf + got);
        std::fclose(f);
    }
    text.push_back('\0');  // strtof/strtod need a terminator
    Cursor cur(text.data(), text.data() + text.size() - 1);

    Banner b;
    int rc = parse_banner(cur, &b);
    if (rc) return rc;
    long M = 0, K = 0, nz = 0;
    if ((rc = parse_size(cur, &M, &K, &nz))) return rc;
    if (!b.coordinate) return fail(SX_ERR_FORMAT, std::string("The input matrix file ") + path + " is not a coordinate file!");
    if (b.complex_) return fail(SX_ERR_FORMAT, "complex matrices are not supported");
    if (M < 0 || K < 0 || nz < 0 || M > INT32_MAX || K > INT32_MAX) return fail(SX_ERR_FORMAT, "bad matrix size");

    struct Entry { int32_t r, c; T v; };
    std::vector<Entry> coo;
    coo.reserve((size_t)nz * (b.symmetric ? 2 : 1));
    for (long e = 0; e < nz; ++e) {
        long r, c;
        T v = T(1);
        if (!cur.integer(&r) || !cur.integer(&c)) return fail(SX_ERR_IO, "entry " + std::to_string(e) + ": missing or malformed indices");
        if (!b.pattern && !cur.real(&v)) return fail(SX_ERR_IO, "entry " + std::to_string(e) + ": missing or malformed value");
        if (is_plus_zero(v)) continue;  // explicit +0 entries are not nonzeros; -0 is kept
        if (r < 1 || c < 1) return fail(SX_ERR_FORMAT, "entry " + std::to_string(e) + ": index below 1");
        if (r > M || c > K) return fail(SX_ERR_FORMAT, "entry " + std::to_string(e) + ": index beyond the declared size");
        coo.push_back({(int32_t)(r - 1), (int32_t)(c - 1), v});
        if (b.symmetric && r != c) {
            if (c > M || r > K) return fail(SX_ERR_FORMAT, "entry " + std::to_string(e) + ": mirrored index beyond the declared size");
            coo.push_back({(int32_t)(c - 1), (int32_t)(r - 1), v});
        }
    }
    const size_t n = coo.size();
    if (n > (size_t)INT32_MAX) return fail(SX_ERR_FORMAT, "more than 2^31-1 nonzeros");

    // stable counting sort by column, then by row  ==  rows ascending, columns
    // ascending inside a row, file order among duplicates
    std::vector<Entry> bycol(n);
    {
        std::vector<size_t> start((size_t)K + 1, 0);
        for (const Entry &en : coo) start[(size_t)en.c + 1]++;
        for (long k = 0; k < K; ++k) start[k + 1] += start[k];
        for (const Entry &en : coo) bycol[start[en.c]++] = en;
    }
    std::vector<Entry>().swap(coo);
    int32_t *rowptr = (int32_t *)std::calloc((size_t)M + 1, sizeof(int32_t));
    int32_t *colidx = (int32_t *)std::malloc(sizeof(int32_t) * std::max<size_t>(n, 1));
    T *val = (T *)std::malloc(sizeof(T) * std::max<size_t>(n, 1));
    if (!rowptr || !colidx || !val) {
        std::free(rowptr); std::free(colidx); std::free(val);
        return fail(SX_ERR_NOMEM, "out of host memory");
    }
    for (const Entry &en : bycol) rowptr[en.r + 1]++;
    for (long i = 0; i < M; ++i) rowptr[i + 1] += rowptr[i];
    {
        std::vector<int32_t> next(rowptr, rowptr + M);
        for (const Entry &en : bycol) {
            const int32_t pos = next[en.r]++;
            colidx[pos] = en.c;
            val[pos] = en.v;
        }
    }
    *M_out = (int)M; *K_out = (int)K; *nnz_out = (int64_t)n;
    *rowptr_out = rowptr; *colidx_out = colidx; *val_out = val;
    return SX_OK;
}

}  // namespace sxhost

// the CUDA translation unit owns sx_last_error(); host-side failures are routed to it
extern "C" void sx_internal_set_error(const char *msg);

extern "C" {

int sx_load_mtx_f32(const char *path, int *M, int *K, int64_t *nnz, int32_t **rowptr, int32_t **colidx, float **val) {
    int rc = sxhost::load<float>(path, M, K, nnz, rowptr, colidx, val);
    if (rc) sx_internal_set_error(sxhost::g_err.c_str());
    return rc;
}

int sx_load_mtx_f64(const char *path, int *M, int *K, int64_t *nnz, int32_t **rowptr, int32_t **colidx, double **val) {
    int rc = sxhost::load<double>(path, M, K, nnz, rowptr, colidx, val);
    if (rc) sx_internal_set_error(sxhost::g_err.c_str());
    return rc;
}

void sx_free(void *ptr) { std::free(ptr); }

int sx_partition_rows(int M, const int32_t *rowptr, int parts, int32_t *bounds) {
    if (M < 0 || parts < 1 || !rowptr || !bounds) {
        sx_internal_set_error("sx_partition_rows: bad argument");
        return SX_ERR_INVALID;
    }
    const int64_t nnz = rowptr[M];
    bounds[0] = 0;
    for (int p = 1; p < parts; ++p) {
        // first row whose starting offset reaches p/parts of the nonzeros; rows are
        // never split, and an all-empty matrix falls back to equal row counts
        int32_t cut;
        if (nnz == 0) cut = (int32_t)((int64_t)M * p / parts);
        else {
            const int64_t target = (nnz * p + parts - 1) / parts;
            cut = (int32_t)(std::lower_bound(rowptr, rowptr + M + 1, target,
                                             [](int32_t a, int64_t t) { return (int64_t)a < t; }) - rowptr);
        }
        bounds[p] = std::min<int32_t>(std::max(cut, bounds[p - 1]), M);
    }
    bounds[parts] = M;
    return SX_OK;
}

}  // extern "C"
