// feht_host.cpp -- the host side of feht above the C ABI, in C++ (the image has no GHC).
//
// Mirrors, function for function, the Haskell host of the reference (paths under /root/reference):
//   src/Table.hs:267-279    parseDataFile            src/Table.hs:160-205  row strings, SNP expansion, replaceChar
//   src/Table.hs:80-115     getMetadataFromFile      src/Table.hs:229-247  filterTable
//   src/Table.hs:251-261    generateCategories       src/Comparison.hs:271-364  getComparisonList, getAllPermutations,
//                                                                               getOva, createComparisonDescription
//   src/Comparison.hs:129-193  formatComparisonResultsAsTable / formatSingleResult (GHC `show` for Int and Double)
//   app/Main.hs:13-31       main
// with Main.hs:26-29 (generateResultMap, applyMultipleTestingCorrection, filter + order) replaced by the
// CUDA library calls of include/feht_cuda.h.  Same argument meaning, same error messages (the reference's
// `error` strings become the returned message).
//
// Where the reference's output order depends on HashMap iteration (comparison blocks, Comparison.hs:129-133;
// rows with equal |ratio|, :160-162) this host is canonical: categories and values ascending, comparisons in
// generation order (pairs, then one-vs-rest), tied rows by marker name.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../include/feht_cuda.h"
#include "../../include/feht_host.h"

namespace {

using Bytes = std::string;

struct HostError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// BS.lines, then BS.filter ('\r' /=) on every line (Table.hs:86,274)
std::vector<Bytes> lines_of(std::string_view bs) {
    std::vector<Bytes> out;
    size_t i = 0;
    while (i < bs.size()) {
        size_t j = bs.find('\n', i);
        if (j == std::string_view::npos) j = bs.size();
        Bytes ln;
        ln.reserve(j - i);
        for (size_t k = i; k < j; ++k)
            if (bs[k] != '\r') ln.push_back(bs[k]);
        out.push_back(std::move(ln));
        i = j + 1;
    }
    return out;
}

// BS.split: "" -> [], "a,,b" -> ["a","","b"], "a," -> ["a",""]
std::vector<Bytes> split(const Bytes &s, char d) {
    std::vector<Bytes> out;
    if (s.empty()) return out;
    size_t i = 0;
    for (;;) {
        size_t j = s.find(d, i);
        if (j == Bytes::npos) { out.emplace_back(s.substr(i)); break; }
        out.emplace_back(s.substr(i, j - i));
        i = j + 1;
    }
    return out;
}

// Table.hs:194-205
char replace_char(char match, char cur) {
    const char n = (cur >= 'a' && cur <= 'z') ? char(cur - 32) : cur;
    if (n == match) return '1';
    if (n == 'A' || n == 'T' || n == 'C' || n == 'G') return '0';
    return n;
}

struct ParsedDataFile {                        // Table.hs:59-62
    std::unordered_map<Bytes, int> name_column_map;
    int n_columns = 0;
    std::vector<Bytes> gene_names;             // GeneVectorMap keys, first occurrence in the file wins
    std::vector<Bytes> gene_vectors;           // the V.Vector Char of each (binary rows; SNP: expanded)
    bool snp = false;
    std::vector<Bytes> site_rows;              // SNP mode: the un-expanded site strings (device-side expansion)
    bool snp_packable = true;                  // no literal '0'/'1' in SNP data
};

ParsedDataFile parse_data_file(char d, bool snp, std::string_view bs) {
    ParsedDataFile pd;
    pd.snp = snp;
    const std::vector<Bytes> lines = lines_of(bs);
    if (lines.empty()) throw HostError("No names present in data file");
    std::vector<Bytes> names = split(lines[0], d);
    if (names.empty()) throw HostError("No genome names present");
    names.erase(names.begin());                // the first header cell is dropped (Table.hs:277)
    for (size_t i = 0; i < names.size(); ++i) pd.name_column_map[names[i]] = int(i);   // M.insert: last wins
    pd.n_columns = int(names.size());
    std::unordered_map<Bytes, size_t> seen;
    static const char *suffix[4] = {"_a", "_t", "_c", "_g"};
    static const char allele[4] = {'A', 'T', 'C', 'G'};
    for (size_t li = 1; li < lines.size(); ++li) {
        const std::vector<Bytes> cells = split(lines[li], d);
        if (cells.empty()) throw HostError("Irrefutable pattern failed for pattern g : gnxs");  // `g:gnxs = xs` on []
        Bytes rest;
        for (size_t c = 1; c < cells.size(); ++c) rest += cells[c];   // BS.concat: cell boundaries vanish
        if (!snp) {
            if (seen.emplace(cells[0], pd.gene_names.size()).second) {
                pd.gene_names.push_back(cells[0]);
                pd.gene_vectors.push_back(std::move(rest));
            }
        } else {
            bool fresh = false;
            for (int a = 0; a < 4; ++a) {
                Bytes nm = cells[0] + suffix[a];
                if (!seen.emplace(nm, pd.gene_names.size()).second) continue;
                fresh = true;
                Bytes v(rest.size(), '\0');
                for (size_t k = 0; k < rest.size(); ++k) v[k] = replace_char(allele[a], rest[k]);
                pd.gene_names.push_back(std::move(nm));
                pd.gene_vectors.push_back(std::move(v));
            }
            if (fresh) {
                for (char ch : rest)
                    if (ch == '0' || ch == '1') pd.snp_packable = false;
                pd.site_rows.push_back(std::move(rest));
            }
            // a duplicated site name whose four rows all exist already adds nothing
            if (pd.site_rows.size() * 4 != pd.gene_names.size()) pd.snp_packable = false;
        }
    }
    return pd;
}

struct MetaRow {                                // one Table entry: GenomeInfo -> MetaHash
    Bytes name;
    int column = 0;
    std::map<Bytes, Bytes> meta;
};
using Table = std::vector<MetaRow>;             // keyed by (name, column): later lines replace earlier ones

Table get_metadata_from_file(char d, const ParsedDataFile &pd, std::string_view bs) {
    if (bs.empty()) throw HostError("File is empty");
    const std::vector<Bytes> lines = lines_of(bs);
    if (lines.size() < 2) throw HostError("Metadata file missing header information, data, or both");
    std::vector<Bytes> cats = split(lines[0], d);
    if (cats.empty()) throw HostError("No metadata categories");
    cats.erase(cats.begin());
    Table t;
    std::unordered_map<Bytes, size_t> index;
    for (size_t li = 1; li < lines.size(); ++li) {
        const std::vector<Bytes> cells = split(lines[li], d);
        if (cells.empty()) throw HostError("Cannot get entry from metadata file");
        if (cells.size() < 2) throw HostError("Metadata line missing name, data, or both");
        auto it = pd.name_column_map.find(cells[0]);
        if (it == pd.name_column_map.end()) throw HostError("name was not found in the data file");
        if (cats.size() != cells.size() - 1) throw HostError("MetaCategories not equal to MeatValues");
        MetaRow r;
        r.name = cells[0];
        r.column = it->second;
        for (size_t c = 0; c < cats.size(); ++c) r.meta[cats[c]] = cells[c + 1];   // M.insert: last wins
        auto ins = index.emplace(r.name, t.size());
        if (ins.second) t.push_back(std::move(r)); else t[ins.first->second] = std::move(r);
    }
    return t;
}

// getListOfColumns . filterTable  (Table.hs:229-247, Comparison.hs:61-68)
std::vector<int32_t> filter_table(const Table &t, const Bytes &mc, bool all_but, const std::vector<Bytes> &mvs) {
    std::vector<int32_t> cols;
    for (const MetaRow &r : t) {
        auto it = r.meta.find(mc);
        if (it == r.meta.end()) continue;
        bool keep;
        if (all_but) {
            keep = true;
            for (const Bytes &v : mvs) keep = keep && (v != it->second);
        } else {
            keep = false;
            for (const Bytes &v : mvs) keep = keep || (v == it->second);
        }
        if (keep) cols.push_back(r.column);
    }
    return cols;
}

std::vector<Bytes> words(const char *s) {
    std::vector<Bytes> out;
    Bytes cur;
    for (const char *p = s; *p; ++p) {
        if (std::isspace(static_cast<unsigned char>(*p))) {
            if (!cur.empty()) { out.push_back(cur); cur.clear(); }
        } else cur.push_back(*p);
    }
    if (!cur.empty()) out.push_back(cur);
    return out;
}

struct GroupCategories { Bytes c1, c2; std::vector<Bytes> v1, v2; };   // Table.hs:251-261
GroupCategories generate_categories(const char *one, const char *two) {
    std::vector<Bytes> o = words(one), t = words(two);
    if (o.empty()) throw HostError("No group one category given");
    if (t.empty()) throw HostError("No group two category given");
    GroupCategories g;
    g.c1 = o[0]; g.v1.assign(o.begin() + 1, o.end());
    g.c2 = t[0]; g.v2.assign(t.begin() + 1, t.end());
    return g;
}

Bytes join(const std::vector<Bytes> &xs, const char *sep) {
    Bytes s;
    for (size_t i = 0; i < xs.size(); ++i) { if (i) s += sep; s += xs[i]; }
    return s;
}

// Comparison.hs:290-308
Bytes description(const Bytes &xc, const std::vector<Bytes> &xs, const Bytes &yc, const std::vector<Bytes> &ys) {
    if (xs.empty()) throw HostError("No group 1 values");
    Bytes g1d = " Group1: " + join(xs, ",") + "\n";
    Bytes g2d = ys.empty() ? " Group2: !" + join(xs, ",")            // no newline (Comparison.hs:307)
                           : " Group2: " + join(ys, ",") + "\n";
    return "Group1 category: " + xc + g1d + "Group2 category: " + yc + g2d;
}

struct PlannedComparison {
    std::vector<int32_t> g1, g2;
    Bytes details;
};
struct PlannedCategory {                        // an "all" category handed to feht_add_category
    Bytes name;
    std::vector<Bytes> values;                  // ascending (Set.toList)
    std::vector<int32_t> value_of_sample;       // per data-file column, -1 = no metadata row
    size_t first_comparison = 0;
};
struct Plan {
    std::vector<PlannedComparison> comps;       // in generation order
    std::vector<PlannedCategory> cats;          // non-empty only for "--one all"
};

// getComparisonList (Comparison.hs:271-285) with getAllPermutations / getOva (:319-344)
Plan get_comparison_list(const Table &t, const GroupCategories &gc, int n_columns) {
    Plan plan;
    if (gc.c1 == "all") {
        std::map<Bytes, std::set<Bytes>> value_sets;
        for (const MetaRow &r : t)
            for (const auto &kv : r.meta) value_sets[kv.first].insert(kv.second);
        for (const auto &cv : value_sets) {
            PlannedCategory pc;
            pc.name = cv.first;
            pc.values.assign(cv.second.begin(), cv.second.end());
            pc.value_of_sample.assign(size_t(n_columns), -1);
            pc.first_comparison = plan.comps.size();
            std::map<Bytes, int> vidx;
            for (size_t i = 0; i < pc.values.size(); ++i) vidx[pc.values[i]] = int(i);
            for (const MetaRow &r : t) {
                auto it = r.meta.find(pc.name);
                if (it != r.meta.end()) pc.value_of_sample[size_t(r.column)] = vidx[it->second];
            }
            const size_t V = pc.values.size();
            for (size_t i = 0; i < V; ++i)
                for (size_t j = i + 1; j < V; ++j)
                    plan.comps.push_back({filter_table(t, pc.name, false, {pc.values[i]}),
                                          filter_table(t, pc.name, false, {pc.values[j]}),
                                          description(pc.name, {pc.values[i]}, pc.name, {pc.values[j]})});
            if (V > 2)
                for (size_t v = 0; v < V; ++v)
                    plan.comps.push_back({filter_table(t, pc.name, false, {pc.values[v]}),
                                          filter_table(t, pc.name, true, {pc.values[v]}),
                                          description(pc.name, {pc.values[v]}, pc.name, {})});
            if (V >= 2) plan.cats.push_back(std::move(pc));
        }
        return plan;
    }
    if (gc.c2 == "all") {
        plan.comps.push_back({filter_table(t, gc.c1, false, gc.v1), filter_table(t, gc.c1, true, gc.v1),
                              description(gc.c1, gc.v1, gc.c1, {})});
        return plan;
    }
    plan.comps.push_back({filter_table(t, gc.c1, false, gc.v1), filter_table(t, gc.c2, false, gc.v2),
                          description(gc.c1, gc.v1, gc.c2, gc.v2)});
    return plan;
}

// GHC `show :: Double -> String` (Numeric.showFloat: shortest round-trip digits; 0.1 <= x < 10^7 positional)
std::string show_double(double x) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x < 0 ? "-Infinity" : "Infinity";
    if (std::signbit(x)) return "-" + show_double(-x);
    if (x == 0) return "0.0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // shortest round trip
    std::string s(buf, res.ptr);
    const size_t epos = s.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; ++i)
        if (s[i] != '.') digits.push_back(s[i]);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int e = std::atoi(s.c_str() + epos + 1) + 1;   // value = 0.d1d2... * 10^e
    if (e >= 0 && e <= 7) {
        if (e == 0) return "0." + digits;
        std::string ds = digits;
        if (int(ds.size()) < e) ds.append(size_t(e) - ds.size(), '0');
        std::string frac = ds.substr(size_t(e));
        return ds.substr(0, size_t(e)) + "." + (frac.empty() ? "0" : frac);
    }
    std::string frac = digits.substr(1);
    return digits.substr(0, 1) + "." + (frac.empty() ? "0" : frac) + "e" + std::to_string(e - 1);
}

const char *kColumnHeader = "Name\tGroupOne (+)\tGroupOne (-)\tGroupTwo (+)\tGroupTwo (-)\tpValue\tRatio\n";

struct Ctx {                                    // RAII over feht_ctx
    feht_ctx *c = nullptr;
    explicit Ctx(int device) {
        if (feht_create(&c, device) != FEHT_OK) throw HostError(std::string("feht_cuda: ") + feht_last_error(nullptr));
    }
    ~Ctx() { if (c) feht_destroy(c); }
    void check(int64_t rc) const {
        if (rc < 0) throw HostError(std::string("feht_cuda: ") + feht_last_error(c));
    }
};

char *dup_out(const std::string &s, int64_t *len) {
    char *p = static_cast<char *>(std::malloc(s.size() + 1));
    if (!p) return nullptr;
    std::memcpy(p, s.data(), s.size());
    p[s.size()] = 0;
    if (len) *len = int64_t(s.size());
    return p;
}

void set_err(char *err, int errlen, const char *msg) {
    if (err && errlen > 0) std::snprintf(err, size_t(errlen), "%s", msg);
}

}  // namespace

extern "C" int feht_host_show_double(double x, char *buf, int buflen) {
    const std::string s = show_double(x);
    if (!buf || buflen <= int(s.size())) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return int(s.size());
}

extern "C" void feht_host_free(char *p) { std::free(p); }

extern "C" int feht_host_plan(const char *metadata, int64_t meta_len, const char *data, int64_t data_len,
                              const char *one, const char *two, char delimiter, int snp_mode, char **out,
                              int64_t *out_len, char *err, int errlen) {
    try {
        const GroupCategories gc = generate_categories(one, two);
        const ParsedDataFile pd = parse_data_file(delimiter, snp_mode != 0, std::string_view(data, size_t(data_len)));
        const Table t = get_metadata_from_file(delimiter, pd, std::string_view(metadata, size_t(meta_len)));
        const Plan plan = get_comparison_list(t, gc, pd.n_columns);
        std::string s = "columns " + std::to_string(pd.n_columns) + " rows " + std::to_string(pd.gene_names.size()) + "\n";
        for (const PlannedComparison &pc : plan.comps) {
            s += "comparison\n";
            s += pc.details;
            if (s.back() != '\n') s += "\n";
            auto cols = [&](std::vector<int32_t> v) {
                std::sort(v.begin(), v.end());
                std::string o;
                for (int32_t x : v) o += std::to_string(x) + " ";
                return o + "\n";
            };
            s += "g1 " + cols(pc.g1);
            s += "g2 " + cols(pc.g2);
        }
        *out = dup_out(s, out_len);
        return *out ? FEHT_OK : FEHT_E_NOMEM;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return FEHT_E_INVALID;
    }
}

extern "C" int feht_host_run(const char *metadata, int64_t meta_len, const char *data, int64_t data_len,
                             const char *one, const char *two, char delimiter, int snp_mode, int correction,
                             double ratio_filter, int device, int engine, char **out, int64_t *out_len,
                             char *err, int errlen) {
    try {
        // ---- Main.hs:16-25: categories, parse, metadata, comparison list
        const GroupCategories gc = generate_categories(one, two);
        const std::string_view whole(data, size_t(data_len));
        // The header line is parsed here; the body goes to the device as text (feht_add_rows_text, N2) unless
        // FEHT_HOST_TEXT=0, marker names repeat (the reference keeps one row per name) or SNP text holds literal
        // '0' / '1' -- then parse_data_file reads the body on the host as before.
        const size_t hdr_end = whole.find('\n');
        const std::string_view header = whole.substr(0, hdr_end == std::string_view::npos ? whole.size() : hdr_end + 1);
        const std::string_view body = hdr_end == std::string_view::npos ? std::string_view() : whole.substr(hdr_end + 1);
        ParsedDataFile pd = parse_data_file(delimiter, snp_mode != 0, header);
        // (metadata and comparisons need the header only: their errors come before anything is uploaded)
        const Table t = get_metadata_from_file(delimiter, pd, std::string_view(metadata, size_t(meta_len)));
        const Plan plan = get_comparison_list(t, gc, pd.n_columns);
        const char *text_env = std::getenv("FEHT_HOST_TEXT");
        bool text_path = !(text_env && std::atoi(text_env) == 0) && pd.n_columns > 0 && !body.empty() && !plan.comps.empty();
        Ctx ctx(device);
        if (text_path) {
            ctx.check(feht_set_samples(ctx.c, pd.n_columns));
            int64_t n_lines = 0;
            const int rc = feht_add_rows_text(ctx.c, body.data(), int64_t(body.size()), delimiter, snp_mode, &n_lines);
            if (rc != FEHT_OK) {
                const std::string msg = feht_last_error(ctx.c);
                if (msg.find("Irrefutable pattern") != std::string::npos)
                    throw HostError("Irrefutable pattern failed for pattern g : gnxs");
                if (msg.find("literal") == std::string::npos) ctx.check(rc);
                text_path = false;
            } else {
                std::vector<int64_t> begin(static_cast<size_t>(n_lines));
                std::vector<int32_t> nlen(begin.size());
                ctx.check(feht_text_lines(ctx.c, begin.data(), nlen.data(), n_lines));
                static const char *sfx[4] = {"_a", "_t", "_c", "_g"};
                pd.gene_names.reserve(begin.size() * (snp_mode ? 4 : 1));
                for (size_t i = 0; i < begin.size(); ++i) {
                    Bytes nm(body.data() + begin[i], size_t(nlen[i]));
                    nm.erase(std::remove(nm.begin(), nm.end(), '\r'), nm.end());
                    if (!snp_mode) pd.gene_names.push_back(std::move(nm));
                    else for (int a = 0; a < 4; ++a) pd.gene_names.push_back(nm + sfx[a]);
                }
            }
        }
        std::vector<int32_t> by_name, name_rank;
        auto rank_names = [&]() -> bool {   // false: a name occurs twice
            const size_t n = pd.gene_names.size();
            by_name.resize(n); name_rank.resize(n);
            for (size_t i = 0; i < n; ++i) by_name[i] = int32_t(i);
            std::sort(by_name.begin(), by_name.end(),
                      [&](int32_t a, int32_t b) { return pd.gene_names[size_t(a)] < pd.gene_names[size_t(b)]; });
            for (size_t r = 0; r < n; ++r) name_rank[size_t(by_name[r])] = int32_t(r);
            for (size_t r = 1; r < n; ++r)
                if (pd.gene_names[size_t(by_name[r])] == pd.gene_names[size_t(by_name[r - 1])]) return false;
            return true;
        };
        if (text_path && !rank_names()) text_path = false;
        if (!text_path) {
            ctx.check(feht_clear_rows(ctx.c));
            pd = parse_data_file(delimiter, snp_mode != 0, whole);
        }

        std::string text;
        const size_t n_rows = pd.gene_names.size();
        if (!plan.comps.empty() && n_rows > 0 && pd.n_columns > 0) {
            // name rank = position of the marker name in byte order (tie-break of the final ordering)
            if (!text_path) rank_names();

            // ---- Main.hs:26-29 on the GPU
            if (!text_path) ctx.check(feht_set_samples(ctx.c, pd.n_columns));
            ctx.check(feht_set_count_engine(ctx.c, engine));
            if (text_path) ctx.check(feht_set_name_rank(ctx.c, name_rank.data(), int64_t(name_rank.size())));
            // SNP sites go up as one character per sample and are expanded on the device, unless the file
            // holds literal '0'/'1' (those count for all four alleles, Table.hs:194-205: only the expanded
            // character rows can express that) or rows are ragged
            bool use_sites = pd.snp && pd.snp_packable;
            for (const Bytes &s : pd.site_rows) use_sites = use_sites && int(s.size()) == pd.n_columns;
            const std::vector<Bytes> &rows = use_sites ? pd.site_rows : pd.gene_vectors;
            std::vector<int64_t> off(rows.size() + 1, 0);
            for (size_t i = 0; i < rows.size(); ++i) off[i + 1] = off[i] + int64_t(rows[i].size());
            std::string flat;
            flat.reserve(size_t(off.back()) + 1);
            for (const Bytes &r : rows) flat += r;
            flat.push_back('\0');
            const uint8_t *chars = reinterpret_cast<const uint8_t *>(flat.data());
            if (text_path) {}   // the rows are on the device already
            else if (use_sites) ctx.check(feht_add_snp_chars(ctx.c, chars, off.data(), int64_t(rows.size()), name_rank.data()));
            else ctx.check(feht_add_rows_chars(ctx.c, chars, off.data(), int64_t(rows.size()), name_rank.data()));
            if (!plan.cats.empty()) {
                for (const PlannedCategory &pc : plan.cats)
                    ctx.check(feht_add_category(ctx.c, pc.value_of_sample.data(), int32_t(pc.values.size())));
            } else {
                for (const PlannedComparison &pc : plan.comps)
                    ctx.check(feht_add_comparison(ctx.c, pc.g1.data(), int64_t(pc.g1.size()), pc.g2.data(),
                                                  int64_t(pc.g2.size())));
            }
            if (feht_num_comparisons(ctx.c) != int64_t(plan.comps.size()))
                throw HostError("internal: planned comparisons and library comparisons differ");
            ctx.check(feht_run(ctx.c, correction, ratio_filter, FEHT_SORT_ABS_RATIO_DESC, 0));

            // ---- Comparison.hs:139-185: one block per comparison with surviving rows
            for (size_t ci = 0; ci < plan.comps.size(); ++ci) {
                const int64_t n = feht_result_count(ctx.c, int64_t(ci));
                ctx.check(n);
                if (n == 0) continue;                                   // Comparison.hs:145-146
                std::vector<int64_t> row(static_cast<size_t>(n));
                std::vector<int32_t> a(row.size()), b(row.size()), c(row.size()), d(row.size());
                std::vector<double> p(row.size()), ratio(row.size());
                ctx.check(feht_result_fetch(ctx.c, int64_t(ci), row.data(), a.data(), b.data(), c.data(), d.data(),
                                            p.data(), ratio.data(), nullptr));
                text += "[#-\n";
                text += plan.comps[ci].details;
                text += "---\n";
                text += kColumnHeader;
                for (size_t i = 0; i < row.size(); ++i) {
                    text += pd.gene_names[size_t(row[i])];
                    text += '\t'; text += std::to_string(a[i]);
                    text += '\t'; text += std::to_string(b[i]);
                    text += '\t'; text += std::to_string(c[i]);
                    text += '\t'; text += std::to_string(d[i]);
                    text += '\t'; text += show_double(p[i]);
                    text += '\t'; text += show_double(ratio[i]);
                    text += '\n';
                }
                text += "-#]\n";
                text += "\n";                                           // BS.putStrLn (Main.hs:30)
            }
        }
        text += "Done\n";                                               // Main.hs:31
        *out = dup_out(text, out_len);
        return *out ? FEHT_OK : FEHT_E_NOMEM;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return FEHT_E_INVALID;
    }
}
