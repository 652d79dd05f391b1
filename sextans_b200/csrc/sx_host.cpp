// sx_host.cpp -- host-side pieces of the drop-in surface that need no GPU:
// the Matrix Market -> CSR loader and the row-block partitioner.
//
// The loader reproduces the OBSERVABLE behaviour of the reference's
// read_suitsparse_matrix(..., CSC) + CSC_2_CSR pair (src/sparse_helper.h:112-259,
// 475-509; banner/size rules of src/mmio.h:254-367) with a different mechanism: the
// file is mapped into memory, the entry region is parsed by one thread per host core
// (the reference calls fscanf per entry), and the (column,row) qsort + CSC->CSR sweep is
// replaced by a bucket-by-row-owner build with a stable per-row sort, which gives the
// same CSR: rows in order, columns ascending within a row.  Entries with equal (row,col)
// keep file order here; the reference's qsort leaves their order unspecified
// (SURVEY.md appendix B).  Measured on 5e6 entries / 132 MB, 8 vCPU: 0.29 s against
// 4.8 s for the reference's loader, same CSR (scripts/bench_loader.py,
// profiles/r01_loader_bench.txt).
//
// Deliberate differences, all on malformed input where the reference has undefined
// or silent behaviour: truncated files, unparsable tokens and indices beyond the
// declared size return an error instead of reusing stale values / writing out of
// bounds.
#include "../../include/sextans_b200.h"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <utility>
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
    const char *pos() const { return p_; }
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

// ---- parallel parsing -------------------------------------------------------------
// for SuiteSparse-sized files the fscanf + qsort loader of the
// reference dwarfs the SpMM.  The entry region is cut at line ends into one chunk per
// host thread; a chunk is accepted only if EVERY non-blank line holds exactly one entry
// (2 tokens for pattern files, 3 otherwise) -- the layout every real .mtx has.  Anything
// else (entries that share or straddle lines, which fscanf tolerates) makes the whole
// file fall back to the serial token parser, so the result never depends on the route.
// Numbers go through std::from_chars, which rounds exactly like strtof/strtod (%f / %lg).
struct Raw { int64_t r, c; };  // 1-based as in the file

template <typename T>
struct ChunkOut {
    std::vector<Raw> idx;
    std::vector<T> val;
    bool regular = true;
    std::string err;
};

static inline const char *skip_blank(const char *p, const char *e) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    return p;
}

template <typename T>
static void parse_chunk(const char *p, const char *e, bool pattern, ChunkOut<T> *out) {
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = nl ? nl : e;
        const char *q = skip_blank(p, le);
        if (q < le) {
            long long r = 0, c = 0;
            auto r1 = std::from_chars(q, le, r);
            if (r1.ec != std::errc()) { out->regular = false; return; }
            q = skip_blank(r1.ptr, le);
            auto r2 = std::from_chars(q, le, c);
            if (r2.ec != std::errc() || r2.ptr == q) { out->regular = false; return; }
            q = skip_blank(r2.ptr, le);
            T v = T(1);
            if (!pattern) {
                if (q < le && *q == '+') ++q;  // from_chars takes no leading plus; strtod does
                auto r3 = std::from_chars(q, le, v);
                if (r3.ec == std::errc::result_out_of_range) { out->regular = false; return; }  // let strtod decide
                if (r3.ec != std::errc() || r3.ptr == q) { out->regular = false; return; }
                q = skip_blank(r3.ptr, le);
            }
            if (q != le) { out->regular = false; return; }  // more tokens on the line
            out->idx.push_back({r, c});
            out->val.push_back(v);
        }
        p = nl ? nl + 1 : e;
    }
}

static unsigned host_threads() {
    if (const char *env = std::getenv("SX_LOADER_THREADS")) {
        const int n = std::atoi(env);
        if (n >= 1) return (unsigned)std::min(n, 256);
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hc ? hc : 1u, 64u));
}

template <typename F>
static void parallel_for(unsigned nthreads, F f) {
    if (nthreads <= 1) { f(0u); return; }
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthreads; ++t) th.emplace_back(f, t);
    f(0u);
    for (auto &x : th) x.join();
}

template <typename T> struct Entry { int32_t r, c; T v; };

// One entry of the file -> 0, 1 or 2 CSR entries (zero drop, 1-based -> 0-based, symmetric
// mirror), bucketed by the thread that owns the row.  Returns 0 or an sx_status with *err.
template <typename T>
static inline int emit_entry(long e, long r, long c, T v, long M, long K, bool symmetric, long rows_per_owner,
                             std::vector<std::vector<Entry<T>>> &bucket, std::string *err) {
    if (is_plus_zero(v)) return SX_OK;  // explicit +0 entries are not nonzeros; -0 is kept
    if (r < 1 || c < 1) { *err = "entry " + std::to_string(e) + ": index below 1"; return SX_ERR_FORMAT; }
    if (r > M || c > K) { *err = "entry " + std::to_string(e) + ": index beyond the declared size"; return SX_ERR_FORMAT; }
    bucket[(size_t)((r - 1) / rows_per_owner)].push_back({(int32_t)(r - 1), (int32_t)(c - 1), v});
    if (symmetric && r != c) {
        if (c > M || r > K) { *err = "entry " + std::to_string(e) + ": mirrored index beyond the declared size"; return SX_ERR_FORMAT; }
        bucket[(size_t)((c - 1) / rows_per_owner)].push_back({(int32_t)(c - 1), (int32_t)(r - 1), v});
    }
    return SX_OK;
}

// buckets[chunk][owner], chunks in file order, owners = contiguous row ranges -> CSR with
// rows ascending, columns ascending inside a row, file order among equal (row, col).
// Every phase runs one thread per owner: row counts, (serial prefix sum), a scatter that
// walks the owner's buckets in chunk order (so it is stable), a stable sort of every row
// that is not already ascending.
template <typename T>
static int buckets_to_csr(const std::vector<std::vector<std::vector<Entry<T>>>> &buckets, long M, unsigned owners,
                          long rows_per_owner, int64_t *nnz_out, int32_t **rowptr_out, int32_t **colidx_out,
                          T **val_out) {
    size_t n = 0;
    for (const auto &ch : buckets)
        for (const auto &bk : ch) n += bk.size();
    if (n > (size_t)INT32_MAX) return fail(SX_ERR_FORMAT, "more than 2^31-1 nonzeros");
    int32_t *rowptr = (int32_t *)std::calloc((size_t)M + 1, sizeof(int32_t));
    int32_t *colidx = (int32_t *)std::malloc(sizeof(int32_t) * std::max<size_t>(n, 1));
    T *val = (T *)std::malloc(sizeof(T) * std::max<size_t>(n, 1));
    if (!rowptr || !colidx || !val) {
        std::free(rowptr); std::free(colidx); std::free(val);
        return fail(SX_ERR_NOMEM, "out of host memory");
    }
    parallel_for(owners, [&](unsigned o) {
        for (const auto &ch : buckets)
            for (const Entry<T> &en : ch[o]) rowptr[en.r + 1]++;
    });
    for (long i = 0; i < M; ++i) rowptr[i + 1] += rowptr[i];
    parallel_for(owners, [&](unsigned o) {
        const long r0 = std::min(M, rows_per_owner * (long)o), r1 = std::min(M, rows_per_owner * (long)(o + 1));
        if (r0 >= r1) return;
        std::vector<int32_t> next(rowptr + r0, rowptr + r1);
        for (const auto &ch : buckets)
            for (const Entry<T> &en : ch[o]) {
                const int32_t pos = next[en.r - r0]++;
                colidx[pos] = en.c;
                val[pos] = en.v;
            }
        std::vector<std::pair<int32_t, T>> tmp;
        for (long r = r0; r < r1; ++r) {
            const int32_t b = rowptr[r], e = rowptr[r + 1];
            bool sorted = true;
            for (int32_t j = b + 1; j < e; ++j)
                if (colidx[j] < colidx[j - 1]) { sorted = false; break; }
            if (sorted) continue;
            tmp.resize((size_t)(e - b));
            for (int32_t j = b; j < e; ++j) tmp[(size_t)(j - b)] = {colidx[j], val[j]};
            std::stable_sort(tmp.begin(), tmp.end(),
                             [](const std::pair<int32_t, T> &x, const std::pair<int32_t, T> &y) { return x.first < y.first; });
            for (int32_t j = b; j < e; ++j) { colidx[j] = tmp[(size_t)(j - b)].first; val[j] = tmp[(size_t)(j - b)].second; }
        }
    });
    *nnz_out = (int64_t)n;
    *rowptr_out = rowptr; *colidx_out = colidx; *val_out = val;
    return SX_OK;
}

// The whole file in memory without a copy; the bytes after the end of the file in its
// last page read as zeros, which is the terminator strtod needs -- unless the size is an
// exact multiple of the page size, in which case the file is copied.
struct FileText {
    const char *data = nullptr;
    size_t size = 0;
    void *map = nullptr;
    size_t map_len = 0;
    std::vector<char> copy;
    ~FileText() { if (map) munmap(map, map_len); }
    int open(const char *path) {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return fail(SX_ERR_IO, std::string("Could not open ") + path);
        struct stat st;
        const long page = sysconf(_SC_PAGESIZE);
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 && st.st_size % page != 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                map = m; map_len = (size_t)st.st_size;
                data = (const char *)m; size = (size_t)st.st_size;
                ::close(fd);
                return SX_OK;
            }
        }
        char buf[1 << 16];
        ssize_t got;
        while ((got = ::read(fd, buf, sizeof buf)) > 0) copy.insert(copy.end(), buf, buf + got);
        ::close(fd);
        size = copy.size();
        copy.push_back('\0');
        data = copy.data();
        return SX_OK;
    }
};

template <typename T>
static int load(const char *path, int *M_out, int *K_out, int64_t *nnz_out, int32_t **rowptr_out,
                int32_t **colidx_out, T **val_out) {
    if (!path || !M_out || !K_out || !nnz_out || !rowptr_out || !colidx_out || !val_out)
        return fail(SX_ERR_INVALID, "null argument");
    *rowptr_out = nullptr; *colidx_out = nullptr; *val_out = nullptr;
    const bool trace = std::getenv("SX_LOADER_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t_start = now();
    FileText text;
    int rc = text.open(path);
    if (rc) return rc;
    Cursor cur(text.data, text.data + text.size);

    Banner b;
    if ((rc = parse_banner(cur, &b))) return rc;
    long M = 0, K = 0, nz = 0;
    if ((rc = parse_size(cur, &M, &K, &nz))) return rc;
    if (!b.coordinate) return fail(SX_ERR_FORMAT, std::string("The input matrix file ") + path + " is not a coordinate file!");
    if (b.complex_) return fail(SX_ERR_FORMAT, "complex matrices are not supported");
    if (M < 0 || K < 0 || nz < 0 || M > INT32_MAX || K > INT32_MAX) return fail(SX_ERR_FORMAT, "bad matrix size");
    const auto t_read = now();

    // ---- entries: parallel when the file is line-regular, serial token parser otherwise ----
    const char *dbeg = cur.pos(), *dend = text.data + text.size;
    const unsigned nthreads = (size_t)(dend - dbeg) < (1u << 20) ? 1u : host_threads();
    const unsigned owners = nthreads;
    const long rows_per_owner = std::max(1L, (M + (long)owners - 1) / (long)owners);
    std::vector<ChunkOut<T>> chunks(nthreads);
    {
        std::vector<const char *> cut(nthreads + 1, dend);
        cut[0] = dbeg;
        for (unsigned t = 1; t < nthreads; ++t) {
            const char *p = dbeg + (size_t)(dend - dbeg) * t / nthreads;
            const char *nl = (const char *)memchr(p, '\n', (size_t)(dend - p));
            cut[t] = nl ? nl + 1 : dend;
        }
        for (unsigned t = 1; t <= nthreads; ++t) cut[t] = std::max(cut[t], cut[t - 1]);
        parallel_for(nthreads, [&](unsigned t) { parse_chunk<T>(cut[t], cut[t + 1], b.pattern, &chunks[t]); });
    }
    const auto t_parse = now();
    bool regular = true;
    size_t found = 0;
    for (const auto &c : chunks) { regular = regular && c.regular; found += c.idx.size(); }

    std::vector<std::vector<std::vector<Entry<T>>>> buckets;
    if (regular && found >= (size_t)nz) {
        // the reference reads exactly nz entries and ignores what follows
        std::vector<long> first(nthreads + 1, 0);
        for (unsigned t = 0; t < nthreads; ++t) first[t + 1] = first[t] + (long)chunks[t].idx.size();
        buckets.assign(nthreads, std::vector<std::vector<Entry<T>>>(owners));
        std::vector<int> status(nthreads, SX_OK);
        std::vector<std::string> errs(nthreads);
        parallel_for(nthreads, [&](unsigned t) {
            const ChunkOut<T> &c = chunks[t];
            const long take = std::max(0L, std::min((long)c.idx.size(), nz - first[t]));
            for (auto &bk : buckets[t]) bk.reserve((size_t)take * (b.symmetric ? 2 : 1) / owners + 16);
            for (long i = 0; i < take; ++i)
                if ((status[t] = emit_entry<T>(first[t] + i, (long)c.idx[(size_t)i].r, (long)c.idx[(size_t)i].c, c.val[(size_t)i], M,
                                               K, b.symmetric, rows_per_owner, buckets[t], &errs[t])))
                    return;
        });
        for (unsigned t = 0; t < nthreads; ++t)  // the first bad entry in file order, as the serial walk would report
            if (status[t]) return fail(status[t], errs[t]);
    } else {
        regular = false;
        buckets.assign(1, std::vector<std::vector<Entry<T>>>(owners));
        std::string err;
        for (long e = 0; e < nz; ++e) {
            long r, c;
            T v = T(1);
            if (!cur.integer(&r) || !cur.integer(&c)) return fail(SX_ERR_IO, "entry " + std::to_string(e) + ": missing or malformed indices");
            if (!b.pattern && !cur.real(&v)) return fail(SX_ERR_IO, "entry " + std::to_string(e) + ": missing or malformed value");
            if ((rc = emit_entry<T>(e, r, c, v, M, K, b.symmetric, rows_per_owner, buckets[0], &err))) return fail(rc, err);
        }
    }
    std::vector<ChunkOut<T>>().swap(chunks);
    const auto t_coo = now();
    if ((rc = buckets_to_csr<T>(buckets, M, owners, rows_per_owner, nnz_out, rowptr_out, colidx_out, val_out))) return rc;
    if (trace)
        std::fprintf(stderr, "sx loader: open+header %.0f ms, parse %.0f ms (%u threads, %s), entries %.0f ms, csr %.0f ms\n",
                     ms(t_start, t_read), ms(t_read, t_parse), nthreads, regular ? "regular" : "serial fallback",
                     ms(t_parse, t_coo), ms(t_coo, now()));
    *M_out = (int)M; *K_out = (int)K;
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

// Column windows (the reference's WINDOW_SIZE partition of A, src/sparse_helper.h:359-371,
// with a window sized for the GPU's L2 instead of the FPGA's on-chip B buffer): window w
// holds the nonzeros with w*W <= col < (w+1)*W as a CSR of its own.  Rows are walked by
// one host thread per contiguous row range; positions follow from prefix sums, so the
// result does not depend on the thread count.  Inside (window, row) the stored order is
// kept, and for rows stored in ascending column order (what the loader produces) walking
// the windows in order visits a row's nonzeros in exactly the stored order.
int sx_split_col_windows(int M, int K, const int32_t *rowptr, const int32_t *colidx, int window_rows, int *nwin_out,
                         int32_t **win_rowptr_out, int64_t **win_base_out, int32_t **order_out, int *ascending_out) {
    if (!nwin_out || !win_rowptr_out || !win_base_out || !order_out) {
        sx_internal_set_error("sx_split_col_windows: null output pointer");
        return SX_ERR_INVALID;
    }
    *nwin_out = 0;
    *win_rowptr_out = nullptr;
    *win_base_out = nullptr;
    *order_out = nullptr;
    if (M < 0 || K < 0 || window_rows < 1 || !rowptr || (rowptr[M] > 0 && !colidx)) {
        sx_internal_set_error("sx_split_col_windows: bad argument");
        return SX_ERR_INVALID;
    }
    const int W = window_rows;
    if (((int64_t)K + W - 1) / W > 4096) {
        sx_internal_set_error("sx_split_col_windows: more than 4096 windows (window_rows too small for K)");
        return SX_ERR_INVALID;
    }
    const int nwin = std::max(1, (int)(((int64_t)K + W - 1) / W));
    const int64_t nnz = rowptr[M];
    const size_t stride = (size_t)M + 1;
    int32_t *wrp = (int32_t *)std::calloc((size_t)nwin * stride, sizeof(int32_t));
    int64_t *base = (int64_t *)std::calloc((size_t)nwin + 1, sizeof(int64_t));
    int32_t *order = (int32_t *)std::malloc(std::max<size_t>((size_t)nnz, 1) * sizeof(int32_t));
    if (!wrp || !base || !order) {
        std::free(wrp); std::free(base); std::free(order);
        sx_internal_set_error("sx_split_col_windows: out of host memory");
        return SX_ERR_NOMEM;
    }
    const unsigned nt = nnz < (1 << 18) ? 1u : sxhost::host_threads();
    auto row_range = [&](unsigned t, int *r0, int *r1) {
        *r0 = (int)((int64_t)M * t / nt);
        *r1 = (int)((int64_t)M * (t + 1) / nt);
    };
    std::vector<int> bad(nt, 0), unsorted(nt, 0);
    // pass 1: wrp[w][r + 1] = nonzeros of row r in window w
    sxhost::parallel_for(nt, [&](unsigned t) {
        int r0, r1;
        row_range(t, &r0, &r1);
        for (int r = r0; r < r1; ++r) {
            int32_t prev = -1;
            for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) {
                const int32_t c = colidx[j];
                if ((uint32_t)c >= (uint32_t)K) { bad[t] = 1; continue; }
                if (c < prev) unsorted[t] = 1;
                prev = c;
                ++wrp[(size_t)(c / W) * stride + r + 1];
            }
        }
    });
    for (unsigned t = 0; t < nt; ++t)
        if (bad[t]) {
            std::free(wrp); std::free(base); std::free(order);
            sx_internal_set_error("sx_split_col_windows: column index out of range");
            return SX_ERR_INVALID;
        }
    // per-window prefix sums (windows are independent)
    sxhost::parallel_for(std::min<unsigned>(nt, (unsigned)nwin), [&](unsigned t) {
        const unsigned step = std::min<unsigned>(nt, (unsigned)nwin);
        for (int w = (int)t; w < nwin; w += (int)step) {
            int32_t *p = wrp + (size_t)w * stride;
            for (int r = 0; r < M; ++r) p[r + 1] += p[r];
        }
    });
    for (int w = 0; w < nwin; ++w) base[w + 1] = base[w] + wrp[(size_t)w * stride + M];
    // pass 2: source position of every entry, window-major
    sxhost::parallel_for(nt, [&](unsigned t) {
        int r0, r1;
        row_range(t, &r0, &r1);
        std::vector<int32_t> fill((size_t)nwin);
        for (int r = r0; r < r1; ++r) {
            for (int w = 0; w < nwin; ++w) fill[w] = wrp[(size_t)w * stride + r];
            for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j) {
                const int w = colidx[j] / W;
                order[base[w] + fill[w]++] = j;
            }
        }
    });
    int asc = 1;
    for (unsigned t = 0; t < nt; ++t) asc &= !unsorted[t];
    if (ascending_out) *ascending_out = asc;
    *nwin_out = nwin;
    *win_rowptr_out = wrp;
    *win_base_out = base;
    *order_out = order;
    return SX_OK;
}

// Plan of the sliding-window kernel (spmm_slide_kernel, variant 4): rows are taken in steps
// of 32; a chain is a run of consecutive steps that one thread block walks with the B rows it
// needs held in a shared-memory ring.  Per step the plan says which NEW rows of B enter the
// ring -- [load_lo, load_hi): everything between the highest row loaded so far and the
// highest column the step touches -- and the step's nonzero range.  ring_rows is the number
// of ring rows that keeps every step's columns resident while the next step's rows are
// already arriving: max over consecutive steps of (highest row loaded for step s+1) -
// (lowest column of step s) + 1.  Chains hold about the same number of nonzeros.
int sx_plan_slide(int M, const int32_t *rowptr, const int32_t *colidx, int nchains_wanted, int *nsteps_out,
                  int32_t **steps_out, int *nchains_out, int32_t **chains_out, int *ring_rows_out,
                  int *max_step_entries_out) {
    if (!nsteps_out || !steps_out || !nchains_out || !chains_out || !ring_rows_out || !max_step_entries_out) {
        sx_internal_set_error("sx_plan_slide: null output pointer");
        return SX_ERR_INVALID;
    }
    *nsteps_out = *nchains_out = *ring_rows_out = *max_step_entries_out = 0;
    *steps_out = *chains_out = nullptr;
    if (M < 0 || nchains_wanted < 1 || !rowptr || (M > 0 && rowptr[M] > 0 && !colidx)) {
        sx_internal_set_error("sx_plan_slide: bad argument");
        return SX_ERR_INVALID;
    }
    const int nsteps = (M + 31) / 32;
    if (nsteps == 0) return SX_OK;
    std::vector<int32_t> cmin((size_t)nsteps), cmax((size_t)nsteps);
    const unsigned nt = rowptr[M] < (1 << 18) ? 1u : sxhost::host_threads();
    sxhost::parallel_for(nt, [&](unsigned t) {
        for (int s = (int)((int64_t)nsteps * t / nt); s < (int)((int64_t)nsteps * (t + 1) / nt); ++s) {
            const int r0 = s * 32, r1 = std::min(M, r0 + 32);
            int32_t lo = INT32_MAX, hi = -1;
            for (int32_t j = rowptr[r0]; j < rowptr[r1]; ++j) { lo = std::min(lo, colidx[j]); hi = std::max(hi, colidx[j]); }
            cmin[s] = lo;  // INT32_MAX / -1: the step has no nonzeros
            cmax[s] = hi;
        }
    });
    // chains: contiguous runs of steps with about nnz / nchains nonzeros each (at least one step)
    const int nchains = std::min(nchains_wanted, nsteps);
    int32_t *chains = (int32_t *)std::malloc((size_t)nchains * 2 * sizeof(int32_t));
    int32_t *steps = (int32_t *)std::malloc((size_t)nsteps * 4 * sizeof(int32_t));
    if (!chains || !steps) {
        std::free(chains); std::free(steps);
        sx_internal_set_error("sx_plan_slide: out of host memory");
        return SX_ERR_NOMEM;
    }
    const int64_t nnz = rowptr[M];
    int s = 0;
    for (int c = 0; c < nchains; ++c) {
        const int begin = s;
        const int must_leave = nchains - 1 - c;  // steps that have to remain for the chains after this one
        const int64_t target = nnz * (c + 1) / nchains;
        ++s;
        while (s < nsteps - must_leave && (c == nchains - 1 || (int64_t)rowptr[std::min(M, s * 32)] < target)) ++s;
        chains[2 * c] = begin;
        chains[2 * c + 1] = s;
    }
    int ring_rows = 1, max_entries = 4;
    for (int c = 0; c < nchains; ++c) {
        const int b = chains[2 * c], e = chains[2 * c + 1];
        int32_t lo0 = INT32_MAX;
        for (int t = b; t < e; ++t) lo0 = std::min(lo0, cmin[t]);
        int32_t hi = (lo0 == INT32_MAX) ? -1 : lo0 - 1;  // highest B row loaded so far
        int32_t prev_cmin = INT32_MAX;                    // lowest column of the previous step
        for (int t = b; t < e; ++t) {
            const int32_t new_hi = std::max(hi, cmax[t]);
            const int r0 = t * 32, r1 = std::min(M, r0 + 32);
            steps[4 * t + 0] = hi + 1;
            steps[4 * t + 1] = new_hi + 1;
            steps[4 * t + 2] = rowptr[r0];
            steps[4 * t + 3] = rowptr[r1];
            if (cmin[t] != INT32_MAX) ring_rows = std::max(ring_rows, new_hi - cmin[t] + 1);
            // while the previous step computes, this step's rows are already arriving
            if (prev_cmin != INT32_MAX) ring_rows = std::max(ring_rows, new_hi - prev_cmin + 1);
            if (t == b && lo0 != INT32_MAX) ring_rows = std::max(ring_rows, new_hi - lo0 + 1);
            max_entries = std::max(max_entries, (rowptr[r1] - (rowptr[r0] & ~3) + 3) & ~3);
            hi = new_hi;
            prev_cmin = cmin[t];
        }
    }
    *nsteps_out = nsteps;
    *steps_out = steps;
    *nchains_out = nchains;
    *chains_out = chains;
    *ring_rows_out = ring_rows;
    *max_step_entries_out = max_entries;
    return SX_OK;
}

// Plan of the edge-list kernel (spmm_edgelist_kernel, variant 5).  The reference cuts A into
// windows of 4096 columns and stores every nonzero of a PE as a packed word whose column field
// is LOCAL to the window (14 bits; src/sparse_helper.h:419-443, src/sextans.cpp:398-402), so
// that the PE indexes its on-chip copy of the B window directly, and it deals rows to PEs so
// that every PE list has the same length (:345-403).  The GPU analogue: a row block (consecutive
// rows, one thread block) stages exactly the B rows its nonzeros touch -- the block's DISTINCT
// columns, in ascending order -- into shared memory, and every nonzero carries a 16-bit index
// into that compacted window; blocks are cut so that they hold about the same number of
// nonzeros.  On FEM-type matrices the compacted window is a third of the contiguous column span
// (nasa4704: 142 distinct columns per 32 rows against a span of 456; pcrystk02: 317 against 918).
// The streams of A are ROW-ALIGNED: row r's entries start at prow[r], a multiple of 8 entries, and are padded to a
// multiple of 8 (prow[r+1] - prow[r] = its length rounded up), so that the kernel fetches a chunk of 8 (local column,
// value) pairs with whole 16-byte shared-memory loads -- 1 + 4 loads instead of 8 + 8 for fp64 -- next to the 8 loads
// of B pieces; the additions of the pad entries are predicated off (the reference's bubbles, src/sparse_helper.h:345-403,
// without their arithmetic).
//   blocks  8 ints per block: {row_begin, nrows, pnz_begin, pnz_end, col_begin, ncols, 0, smem_bytes}, pnz in PADDED
//           coordinates: prow[row_begin], prow[row_begin + nrows]
//   cols    the blocks' column lists, back to back, each starting at a multiple of 4 entries
//           (16 bytes: the list travels to shared memory by TMA); pad entries repeat the last column
//   prow    M + 1 padded row starts
//   lcol    prow[M] 16-bit local column indices: entry k of row r at prow[r] + k, pad entries 0 (a valid local
//           column whenever the row has entries)
// A block is closed when it holds max_rows rows, when its nonzeros reach nnz_target (0: no such
// limit; compared at the row that brings it closest), or when the next row would not fit the
// shared-memory budget; a single row that does not fit makes the matrix unplannable (*nblocks =
// 0, SX_OK).  Cuts restart every 4096 rows, so the plan does not depend on the thread count.
int sx_plan_edge_lists(int M, int K, const int32_t *rowptr, const int32_t *colidx, int row_bytes, int elem_bytes,
                       int max_rows, int64_t nnz_target, int smem_budget, int *nblocks_out, int32_t **blocks_out,
                       int64_t *ncols_out, int32_t **cols_out, uint16_t **lcol_out, int64_t *total_cols_out,
                       int *max_smem_out, int32_t **prow_out) {
    if (!nblocks_out || !blocks_out || !ncols_out || !cols_out || !lcol_out || !total_cols_out || !max_smem_out || !prow_out) {
        sx_internal_set_error("sx_plan_edge_lists: null output pointer");
        return SX_ERR_INVALID;
    }
    *nblocks_out = *max_smem_out = 0;
    *ncols_out = *total_cols_out = 0;
    *blocks_out = *cols_out = *prow_out = nullptr;
    *lcol_out = nullptr;
    if (M < 0 || K < 0 || !rowptr || (M > 0 && rowptr[M] > 0 && !colidx) || row_bytes < 16 || row_bytes % 16 ||
        (elem_bytes != 4 && elem_bytes != 8) || smem_budget < 1024 || max_rows < 1 || max_rows > 4096 || nnz_target < 0) {
        sx_internal_set_error("sx_plan_edge_lists: bad argument");
        return SX_ERR_INVALID;
    }
    if (M == 0) return SX_OK;
    constexpr int SG = 4096;  // rows per super-group: cuts restart here
    const int64_t nnz = rowptr[M];
    const int ngroups = (M + SG - 1) / SG;
    // padded row starts
    int32_t *prow = (int32_t *)std::malloc(((size_t)M + 1) * sizeof(int32_t));
    if (!prow) { sx_internal_set_error("sx_plan_edge_lists: out of host memory"); return SX_ERR_NOMEM; }
    {
        int64_t at = 0;
        for (int r = 0; r < M; ++r) {
            prow[r] = (int32_t)at;
            at += ((int64_t)(rowptr[r + 1] - rowptr[r]) + 7) & ~(int64_t)7;
            if (at > INT32_MAX - 8) { std::free(prow); return SX_OK; }  // not plannable: the caller keeps its other kernels
        }
        prow[M] = (int32_t)at;
    }
    const int64_t pnz = prow[M];
    // shared memory of a block: window | values | local columns | column list | padded row starts | row ends
    // (pb, pe in padded coordinates: both streams are whole 16-byte units)
    auto smem_of = [&](int ncols, int nrows, int pb, int pe) -> int64_t {
        const int64_t na = pe - pb;
        return (int64_t)ncols * row_bytes + na * elem_bytes + na * 2 + (int64_t)((ncols + 3) & ~3) * 4 + 2 * (int64_t)((nrows + 4) & ~3) * 4;
    };
    struct Part { std::vector<int32_t> blocks, cols; int64_t total = 0; int max_smem = 0; bool ok = true; };
    const unsigned nt = (unsigned)std::min<int64_t>(nnz < (1 << 18) ? 1 : std::min(sxhost::host_threads(), 16u), ngroups);  // 6 bytes x K of scratch per thread
    std::vector<Part> parts(nt);
    uint16_t *lcol = (uint16_t *)std::calloc(std::max<size_t>((size_t)pnz, 8), sizeof(uint16_t));  // pad entries: 0
    if (!lcol) { std::free(prow); sx_internal_set_error("sx_plan_edge_lists: out of host memory"); return SX_ERR_NOMEM; }
    sxhost::parallel_for(nt, [&](unsigned t) {
        Part &P = parts[t];
        std::vector<int32_t> stamp((size_t)K, -1), cols;
        std::vector<uint16_t> local((size_t)K, 0);
        int32_t tag = 0;
        const int g0 = (int)((int64_t)ngroups * t / nt), g1 = (int)((int64_t)ngroups * (t + 1) / nt);
        for (int g = g0; g < g1 && P.ok; ++g) {
            const int gend = (int)std::min<int64_t>(M, (int64_t)g * SG + SG);
            int r = g * SG;
            while (r < gend) {
                // greedy: rows r.. while the block still fits and is short of its share of nonzeros
                const int rb = r;
                const int jb = rowptr[rb];
                int ncols = 0;
                ++tag;
                cols.clear();
                while (r < gend && r - rb < max_rows) {
                    const int64_t have = rowptr[r] - jb, with = rowptr[r + 1] - jb;
                    if (nnz_target > 0 && r > rb && with - nnz_target > nnz_target - have) break;  // closer to the target without this row
                    int fresh = 0;
                    for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j)
                        if (stamp[colidx[j]] != tag) { stamp[colidx[j]] = tag; cols.push_back(colidx[j]); ++fresh; }
                    if (smem_of(ncols + fresh, r - rb + 1, prow[rb], prow[r + 1]) > smem_budget || ncols + fresh > 65535) {
                        if (r == rb) { P.ok = false; break; }  // one row alone does not fit
                        for (int k = 0; k < fresh; ++k) { stamp[cols.back()] = -1; cols.pop_back(); }
                        break;
                    }
                    ncols += fresh;
                    ++r;
                }
                if (!P.ok) break;
                std::sort(cols.begin(), cols.end());
                for (int i = 0; i < ncols; ++i) local[cols[i]] = (uint16_t)i;
                for (int rr = rb; rr < r; ++rr)
                    for (int32_t j = rowptr[rr]; j < rowptr[rr + 1]; ++j) lcol[prow[rr] + (j - rowptr[rr])] = local[colidx[j]];
                const int sm = (int)smem_of(ncols, r - rb, prow[rb], prow[r]);
                P.blocks.insert(P.blocks.end(), {rb, r - rb, prow[rb], prow[r], (int32_t)P.cols.size(), ncols, 0, sm});
                P.cols.insert(P.cols.end(), cols.begin(), cols.end());
                while (P.cols.size() % 4) P.cols.push_back(ncols ? cols.back() : 0);
                P.total += ncols;
                P.max_smem = std::max(P.max_smem, sm);
            }
        }
    });
    bool ok = true;
    size_t nb = 0, nc = 0;
    for (const Part &P : parts) { ok = ok && P.ok; nb += P.blocks.size() / 8; nc += P.cols.size(); }
    if (!ok || nb > (size_t)INT32_MAX || nc > (size_t)INT32_MAX) {
        std::free(lcol);
        std::free(prow);
        return SX_OK;  // not plannable with this budget: the caller keeps its other kernels
    }
    int32_t *blocks = (int32_t *)std::malloc(std::max<size_t>(nb, 1) * 8 * sizeof(int32_t));
    int32_t *cols = (int32_t *)std::malloc(std::max<size_t>(nc, 4) * sizeof(int32_t));
    if (!blocks || !cols) {
        std::free(blocks); std::free(cols); std::free(lcol); std::free(prow);
        sx_internal_set_error("sx_plan_edge_lists: out of host memory");
        return SX_ERR_NOMEM;
    }
    size_t bo = 0, co = 0;
    int64_t total = 0;
    int max_smem = 0;
    for (const Part &P : parts) {
        for (size_t i = 0; i < P.blocks.size(); i += 8) {
            std::copy(P.blocks.begin() + i, P.blocks.begin() + i + 8, blocks + (bo + i));
            blocks[bo + i + 4] += (int32_t)co;  // column-list offsets become global
        }
        std::copy(P.cols.begin(), P.cols.end(), cols + co);
        bo += P.blocks.size();
        co += P.cols.size();
        total += P.total;
        max_smem = std::max(max_smem, P.max_smem);
    }
    *nblocks_out = (int)nb;
    *blocks_out = blocks;
    *ncols_out = (int64_t)nc;
    *cols_out = cols;
    *lcol_out = lcol;
    *prow_out = prow;
    *total_cols_out = total;
    *max_smem_out = max_smem;
    return SX_OK;
}

}  // extern "C"
