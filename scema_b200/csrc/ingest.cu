// Batch ingest of strain histories from files (host side, multi-threaded) — SURVEY.md §8(f)-2.
//
//  * strain_<ID> text files: what Strain6D::from_file does one history at a time through
//    `ifstream >> double` (reference headers/strain2spline.h:112-134), as the two command lines
//    drive it (clustering/mpi_comparison_test.cc:67-88, clustering/compare_all_histories.cc:41-57):
//    here all files are parsed concurrently into ONE ragged batch [sum L_i][6] + offsets + ids,
//    which is exactly the argument list of scema_set_histories.
//  * pr_<rank>.lhistory.csv: the per-rank history log written by FEProblem::output_lhistory
//    (reference headers/FE_problem.h:1985-2045): rows `timestep,time,qpid,cell,qpoint,material,
//    strain_00,strain_01,strain_02,strain_11,strain_12,strain_22,updstrain_..x6,stress_..x6`.
//    Rows are grouped by qpid in file order and the tensor components reordered from the log's
//    00,01,02,11,12,22 to Strain6D's xx,yy,zz,xy,xz,yz (FE_problem.h:1092-1098).
//
// Number syntax. `istream >> double` (libstdc++ num_get) collects characters by a fixed grammar —
// [sign] digits [. digits] [e|E [sign] digits], at most one point, exponent only after a mantissa
// digit — hands them to strtod and fails unless strtod consumes all of them; a failure ends the
// read loop of from_file, and a partially read line is dropped. scan_number() restates that grammar;
// the value comes from an exact fast path (<= 15 significant digits and |exponent| <= 22: one
// correctly rounded multiplication or division, Clinger 1990) or from strtod itself.
#include "common.cuh"
#include "../../include/scema_ingest.h"

#include <atomic>
#include <cerrno>
#include <cmath>
#include <dirent.h>
#include <map>
#include <thread>

struct scema_batch {
    std::vector<double> steps;      // [offsets.back()][6]
    std::vector<uint64_t> offsets;  // [n+1]
    std::vector<uint32_t> ids;      // [n]
    std::vector<std::string> names; // [n] source (file name, or "qpid <id>")
};

namespace {

thread_local std::string g_ingest_error;

int ingest_fail(int code, const std::string &msg)
{
    g_ingest_error = msg;
    return code;
}

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// One `in >> double` at *pp (leading whitespace skipped). Returns false when the extraction fails
// (end of data, malformed token, or overflow to infinity — libstdc++ sets failbit for those).
// On success *pp is left after the token.
bool scan_number(const char **pp, const char *end, double *out)
{
    const char *p = *pp;
    while (p < end && is_space(*p)) p++;
    if (p >= end) return false;
    const char *tok = p;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; p++; }
    uint64_t mant = 0;       // first 19 significant digits
    int sig = 0;             // significant digits seen (leading zeros excluded)
    int dropped = 0;         // integer-part digits beyond the 19 kept
    int frac_kept = 0;       // fractional digits that went into mant
    bool any_digit = false, point = false, simple = true;
    while (p < end) {
        const char ch = *p;
        if (is_digit(ch)) {
            any_digit = true;
            if (sig > 0 || ch != '0') {
                if (sig < 19) { mant = mant * 10 + (uint64_t)(ch - '0'); if (point) frac_kept++; }
                else { if (!point) dropped++; if (ch != '0') simple = false; }
                sig++;
            } else if (point) {
                frac_kept++;  // zero right after the point, before any significant digit
            }
            p++;
        } else if (ch == '.' && !point) {
            point = true;
            p++;
        } else {
            break;
        }
    }
    int exp10 = 0;
    bool exp_ok = true;
    if (p < end && (*p == 'e' || *p == 'E') && any_digit) {
        p++;
        bool eneg = false;
        if (p < end && (*p == '+' || *p == '-')) { eneg = *p == '-'; p++; }
        if (p >= end || !is_digit(*p)) exp_ok = false;  // "1e", "1e+": strtod stops before the 'e' -> failbit
        int ev = 0;
        while (p < end && is_digit(*p)) { if (ev < 100000) ev = ev * 10 + (*p - '0'); p++; }
        exp10 = eneg ? -ev : ev;
    }
    *pp = p;
    if (!any_digit || !exp_ok) return false;
    // value = mant * 10^(exp10 + dropped - frac_kept)
    const long e = (long)exp10 + dropped - frac_kept;
    double v;
    if (mant == 0 && simple) {
        v = 0.0;
    } else if (simple && sig <= 15 && e >= -22 && e <= 22) {
        v = e >= 0 ? (double)mant * kPow10[e] : (double)mant / kPow10[-e];
    } else if (simple && sig <= 15 && e > 22 && e <= 22 + 15 - sig) {
        v = ((double)mant * kPow10[e - 22]) * kPow10[22];  // mant * 10^(e-22) is still an exact integer < 10^15
    } else {
        char small[128];
        const size_t len = (size_t)(p - tok);
        std::string big;
        const char *z;
        if (len < sizeof small) { memcpy(small, tok, len); small[len] = 0; z = small; }
        else { big.assign(tok, len); z = big.c_str(); }
        char *stop = nullptr;
        v = strtod(z, &stop);
        if (stop == z || *stop != 0) return false;
        if (std::isinf(v)) return false;  // out of range: failbit
        *out = v;
        return true;
    }
    *out = neg ? -v : v;
    return true;
}

// `while (in >> xx >> yy >> zz >> xy >> xz >> yz)` over a whole file image
void parse_strain_text(const char *buf, size_t len, std::vector<double> &out)
{
    const char *p = buf, *end = buf + len;
    double s[6];
    while (true) {
        int k = 0;
        for (; k < 6; k++)
            if (!scan_number(&p, end, &s[k])) break;
        if (k < 6) break;
        out.insert(out.end(), s, s + 6);
    }
}

bool slurp(const std::string &path, std::string &data)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char chunk[1 << 16];
    size_t got;
    data.clear();
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) data.append(chunk, got);
    fclose(f);
    return true;
}

unsigned id_from_name(std::string name)  // atoi of the name without its first "strain_" (mpi_comparison_test.cc:85-87)
{
    size_t pos = name.find("strain_");
    if (pos != std::string::npos) name.erase(pos, 7);
    return (unsigned)atoi(name.c_str());
}

int read_files(const std::vector<std::string> &paths, const std::vector<std::string> &names, const std::vector<uint32_t> &ids,
               uint32_t n_threads, scema_batch **out)
{
    const size_t n = paths.size();
    std::vector<std::vector<double> > parts(n);
    std::atomic<size_t> next(0);
    std::atomic<long> failed(-1);
    if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
    n_threads = (uint32_t)std::min<size_t>(n_threads, std::max<size_t>(n, 1));
    auto work = [&]() {
        std::string data;
        while (true) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            if (!slurp(paths[i], data)) {
                long expect = -1;
                failed.compare_exchange_strong(expect, (long)i);
                continue;
            }
            parse_strain_text(data.data(), data.size(), parts[i]);
        }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < n_threads; t++) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (failed.load() >= 0)
        return ingest_fail(SCEMA_ERR_IO, "Could not open " + paths[(size_t)failed.load()] + " for reading.");
    scema_batch *b = new (std::nothrow) scema_batch();
    if (!b) return ingest_fail(SCEMA_ERR_NOMEM, "out of memory");
    b->offsets.resize(n + 1, 0);
    for (size_t i = 0; i < n; i++) b->offsets[i + 1] = b->offsets[i] + parts[i].size() / 6;
    b->steps.resize((size_t)b->offsets[n] * 6);
    for (size_t i = 0; i < n; i++)
        if (!parts[i].empty()) memcpy(b->steps.data() + b->offsets[i] * 6, parts[i].data(), parts[i].size() * sizeof(double));
    b->ids = ids;
    b->names = names;
    *out = b;
    return SCEMA_OK;
}

// split a CSV line at commas (fields are plain numbers / identifiers in this log; no quoting)
void split_csv(const std::string &line, std::vector<std::string> &f)
{
    f.clear();
    size_t a = 0;
    while (true) {
        size_t c = line.find(',', a);
        std::string cell = line.substr(a, c == std::string::npos ? std::string::npos : c - a);
        size_t i = 0, j = cell.size();
        while (i < j && is_space(cell[i])) i++;
        while (j > i && is_space(cell[j - 1])) j--;
        f.push_back(cell.substr(i, j - i));
        if (c == std::string::npos) break;
        a = c + 1;
    }
}

}  // namespace

extern "C" {

const char *scema_ingest_last_error(void) { return g_ingest_error.c_str(); }

int scema_batch_read_files(const char *const *paths, const uint32_t *ids, uint64_t n, uint32_t n_threads, scema_batch **out)
{
    if (!out || (n && !paths)) return ingest_fail(SCEMA_ERR_INVALID, "read_files: null pointer");
    g_ingest_error.clear();
    std::vector<std::string> p(n), nm(n);
    std::vector<uint32_t> id(n);
    for (uint64_t i = 0; i < n; i++) {
        p[i] = paths[i];
        size_t slash = p[i].rfind('/');
        nm[i] = slash == std::string::npos ? p[i] : p[i].substr(slash + 1);
        id[i] = ids ? ids[i] : id_from_name(nm[i]);
    }
    return read_files(p, nm, id, n_threads, out);
}

int scema_batch_read_dir(const char *strain_directory, uint32_t n_threads, scema_batch **out)
{
    if (!out || !strain_directory) return ingest_fail(SCEMA_ERR_INVALID, "read_dir: null pointer");
    g_ingest_error.clear();
    DIR *d = opendir(strain_directory);
    if (!d) return ingest_fail(SCEMA_ERR_IO, std::string("Could not open directory ") + strain_directory);
    std::vector<std::string> p, nm;
    std::vector<uint32_t> id;
    const std::string dir = strain_directory;  // concatenated as given (the reference needs the trailing '/')
    while (struct dirent *e = readdir(d)) {
        const std::string name = e->d_name;
        if (name.compare(0, 7, "strain_") != 0) continue;
        nm.push_back(name);
        p.push_back(dir + name);
        id.push_back(id_from_name(name));
    }
    closedir(d);
    return read_files(p, nm, id, n_threads, out);
}

int scema_batch_from_lhistory(const char *const *csv_paths, uint64_t n_files, const char *column_prefix, scema_batch **out)
{
    if (!out || (n_files && !csv_paths)) return ingest_fail(SCEMA_ERR_INVALID, "from_lhistory: null pointer");
    g_ingest_error.clear();
    const std::string prefix = column_prefix && *column_prefix ? column_prefix : "strain";
    static const char *suffix[6] = {"_00", "_11", "_22", "_01", "_02", "_12"};  // xx yy zz xy xz yz
    std::map<uint32_t, std::vector<double> > by_qp;  // ascending qpid
    std::string data, line;
    std::vector<std::string> f;
    for (uint64_t q = 0; q < n_files; q++) {
        if (!slurp(csv_paths[q], data))
            return ingest_fail(SCEMA_ERR_IO, std::string("Could not open ") + csv_paths[q] + " for reading.");
        size_t pos = 0;
        int col_qpid = -1, col[6] = {-1, -1, -1, -1, -1, -1};
        size_t n_cols = 0;
        uint64_t line_no = 0;
        while (pos < data.size()) {
            size_t nl = data.find('\n', pos);
            line.assign(data, pos, nl == std::string::npos ? std::string::npos : nl - pos);
            pos = nl == std::string::npos ? data.size() : nl + 1;
            line_no++;
            if (line.empty() || (line.size() == 1 && line[0] == '\r')) continue;
            split_csv(line, f);
            if (col_qpid < 0 || f[0] == "timestep") {  // header (repeated when a run was restarted into the same file)
                if (f[0] != "timestep")
                    return ingest_fail(SCEMA_ERR_INVALID, std::string(csv_paths[q]) + ": missing header line");
                col_qpid = -1;
                for (int k = 0; k < 6; k++) col[k] = -1;
                for (size_t c = 0; c < f.size(); c++) {
                    if (f[c] == "qpid") col_qpid = (int)c;
                    for (int k = 0; k < 6; k++)
                        if (f[c] == prefix + suffix[k]) col[k] = (int)c;
                }
                n_cols = f.size();
                if (col_qpid < 0) return ingest_fail(SCEMA_ERR_INVALID, std::string(csv_paths[q]) + ": no qpid column");
                for (int k = 0; k < 6; k++)
                    if (col[k] < 0)
                        return ingest_fail(SCEMA_ERR_INVALID, std::string(csv_paths[q]) + ": no column " + prefix + suffix[k]);
                continue;
            }
            if (f.size() != n_cols)
                return ingest_fail(SCEMA_ERR_INVALID, std::string(csv_paths[q]) + ": line " + std::to_string(line_no) +
                                                          " has " + std::to_string(f.size()) + " fields, header has " +
                                                          std::to_string(n_cols));
            const uint32_t qp = (uint32_t)strtoul(f[col_qpid].c_str(), nullptr, 10);
            std::vector<double> &h = by_qp[qp];
            for (int k = 0; k < 6; k++) {
                const std::string &cell = f[col[k]];
                const char *p = cell.data();
                double v;
                if (!scan_number(&p, cell.data() + cell.size(), &v) || p != cell.data() + cell.size())
                    return ingest_fail(SCEMA_ERR_INVALID, std::string(csv_paths[q]) + ": line " + std::to_string(line_no) +
                                                              ": bad number '" + cell + "'");
                h.push_back(v);
            }
        }
    }
    scema_batch *b = new (std::nothrow) scema_batch();
    if (!b) return ingest_fail(SCEMA_ERR_NOMEM, "out of memory");
    b->offsets.push_back(0);
    for (auto &kv : by_qp) {
        b->ids.push_back(kv.first);
        b->names.push_back("qpid " + std::to_string(kv.first));
        b->steps.insert(b->steps.end(), kv.second.begin(), kv.second.end());
        b->offsets.push_back(b->offsets.back() + kv.second.size() / 6);
    }
    *out = b;
    return SCEMA_OK;
}

uint64_t scema_batch_count(const scema_batch *b) { return b ? b->ids.size() : 0; }
uint64_t scema_batch_total_steps(const scema_batch *b) { return b ? b->offsets.back() : 0; }
const double *scema_batch_steps(const scema_batch *b) { return b ? b->steps.data() : nullptr; }
const uint64_t *scema_batch_offsets(const scema_batch *b) { return b ? b->offsets.data() : nullptr; }
const uint32_t *scema_batch_ids(const scema_batch *b) { return b ? b->ids.data() : nullptr; }
const char *scema_batch_name(const scema_batch *b, uint64_t i) { return b && i < b->names.size() ? b->names[i].c_str() : ""; }

int scema_batch_write_strain_files(const scema_batch *b, const char *out_directory)
{
    if (!b || !out_directory) return ingest_fail(SCEMA_ERR_INVALID, "write_strain_files: null pointer");
    g_ingest_error.clear();
    std::string dir = out_directory;
    if (!dir.empty() && dir.back() != '/') dir += '/';
    std::string text;
    char line[256];
    for (size_t i = 0; i < b->ids.size(); i++) {
        const std::string path = dir + "strain_" + std::to_string(b->ids[i]);
        FILE *f = fopen(path.c_str(), "w");
        if (!f) return ingest_fail(SCEMA_ERR_IO, "Could not open " + path + " for writing.");
        text.clear();
        for (uint64_t s = b->offsets[i]; s < b->offsets[i + 1]; s++) {
            const double *v = b->steps.data() + s * 6;
            // 17 significant digits: every double survives the text round trip
            int len = snprintf(line, sizeof line, "%.17g %.17g %.17g %.17g %.17g %.17g\n", v[0], v[1], v[2], v[3], v[4], v[5]);
            text.append(line, len);
        }
        const bool ok = text.empty() || fwrite(text.data(), 1, text.size(), f) == text.size();
        fclose(f);
        if (!ok) return ingest_fail(SCEMA_ERR_IO, "short write to " + path);
    }
    return SCEMA_OK;
}

int scema_set_histories_from_batch(scema_ctx *ctx, const scema_batch *b)
{
    if (!ctx || !b) return SCEMA_ERR_INVALID;
    return scema_set_histories(ctx, b->steps.data(), 0, b->offsets.data(), b->ids.data(), b->ids.size());
}

void scema_batch_free(scema_batch *b) { delete b; }

}  // extern "C"
