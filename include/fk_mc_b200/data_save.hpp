// Host-side error analysis and output of the Monte Carlo series, header-only C++17: the C++ twin of fk_mc_b200/stats.py and
// fk_mc_b200/h5out.py, for callers that stay in C++ like the reference does.
//
// Mirrors include/fk_mc/binning.hpp:89-171 (calc_stats, bin<D>, accumulate_binning, calc_cor_length),
// include/fk_mc/jackknife.hpp:50-82 (jack, accumulate_jackknife), prog/data_save.hpp:108-156 (estimate_bin, save_bin_data /
// save_binning) and prog/data_save.hxx:33-199 (save_all_data: /parameters, /mc_data/*, energy / d2energy / c_energy / cv
// statistics).  A bin-stats row is (n, mean, variance, stderr) like the reference's bin_stats_t.
//
// The HDF5 container is emitted directly (there is no libhdf5 in this image): version-0 superblock, version-1 object headers,
// symbol-table groups (one level-0 B-tree node + local heap + symbol-table nodes of 2K entries) and contiguous little-endian
// datasets -- the structures libhdf5 writes with its default ("earliest") format bounds, so h5py reads the files as they are
// (scripts/parse/parse_thermod.py:48-49 does `(nbins, value, disp, error) = h5["stats"][obs]`).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

namespace fk {

using bin_stats_t = std::array<double, 4>;  // n, mean, unbiased variance, sqrt(variance / n)

namespace binning {

constexpr int max_bin_depth_compiled = 15;  // BINNING_RANGE, include/fk_mc/binning.hpp:18

inline bin_stats_t calc_stats(const std::vector<double>& x) {  // binning.hpp:89-96
    const size_t n = x.size();
    double mean = 0.0;
    for (double v : x) mean += v;
    mean /= double(n);
    double var = 0.0;
    for (double v : x) var += (v - mean) * (v - mean);
    var = n > 1 ? var / double(n - 1) : std::nan("");
    return {double(n), mean, var, n > 1 ? std::sqrt(var / double(n)) : std::nan("")};
}

// averages of 2^depth consecutive samples, incomplete tail dropped (binned_iterator, binning.hpp:26-44)
inline std::vector<double> bin_series(const std::vector<double>& x, int depth) {
    const size_t step = size_t(1) << depth;
    if (step > x.size()) throw std::logic_error("Can't bin with binning step(" + std::to_string(step) + ")> container size (" + std::to_string(x.size()) + ")");
    const size_t n = x.size() / step;
    std::vector<double> out(n);
    for (size_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (size_t j = 0; j < step; ++j) s += x[i * step + j];
        out[i] = s / double(step);
    }
    return out;
}

inline std::vector<bin_stats_t> accumulate_binning(const std::vector<double>& x, int max_depth) {  // binning.hpp:100-128
    if (max_depth > max_bin_depth_compiled) throw std::logic_error("bin_depth =" + std::to_string(max_depth) + "> compiled bin size");
    std::vector<bin_stats_t> rows;
    for (int d = 0; d <= max_depth; ++d) rows.push_back(calc_stats(bin_series(x, d)));
    return rows;
}

inline std::vector<double> calc_cor_length(const std::vector<bin_stats_t>& rows) {  // binning.hpp:163-171
    std::vector<double> out;
    for (size_t i = 0; i < rows.size(); ++i) out.push_back(0.5 * (std::ldexp(1.0, int(i)) * rows[i][2] / rows[0][2] - 1.0));
    return out;
}

}  // namespace binning

namespace jackknife {

// jackknife of F(<x_1>, <x_2>, ...) over binned series (jackknife.hpp:50-82)
inline bin_stats_t jack(const std::function<double(const std::vector<double>&)>& F, const std::vector<std::vector<double>>& series, int depth) {
    std::vector<std::vector<double>> data;
    for (auto& s : series) data.push_back(binning::bin_series(s, depth));
    const size_t n = data[0].size(), k = data.size();
    std::vector<double> means(k, 0.0), loo(k);
    for (size_t i = 0; i < k; ++i) {
        for (double v : data[i]) means[i] += v;
        means[i] /= double(n);
    }
    const double u0 = F(means);
    std::vector<double> u(n);
    for (size_t j = 0; j < n; ++j) {
        for (size_t i = 0; i < k; ++i) loo[i] = (double(n) * means[i] - data[i][j]) / double(n - 1);
        u[j] = F(loo);
    }
    const bin_stats_t st = binning::calc_stats(u);
    const double u_avg = u0 - double(n - 1) * (st[1] - u0), du = double(n - 1) * st[3];
    return {double(n), u_avg, du * du * double(n), du};
}

inline std::vector<bin_stats_t> accumulate_jackknife(const std::function<double(const std::vector<double>&)>& F,
                                                     const std::vector<std::vector<double>>& series, int max_depth) {
    std::vector<bin_stats_t> rows;
    for (int d = 0; d <= max_depth; ++d) rows.push_back(jack(F, series, d));
    return rows;
}

}  // namespace jackknife

// index of the bin level where the error bar has saturated (prog/data_save.hpp:108-122)
inline size_t estimate_bin(const std::vector<bin_stats_t>& rows) {
    double rel_error = 1.0;
    bool f = true;
    size_t ind = rows.size() - 1;
    while (f && ind > 0) {
        const double cur = std::abs(rows[ind - 1][3] / rows[ind][3] - 1.0);
        f = cur < 0.05 && cur < rel_error;
        if (f) { rel_error = cur; --ind; }
    }
    return ind;
}

// deepest level that still leaves `min_bins` bins, capped at the reference's compile-time depth
inline int max_bin_depth(size_t n_samples, size_t min_bins = 4) {
    int d = 0;
    while (d < binning::max_bin_depth_compiled && (n_samples >> (d + 1)) >= min_bins) ++d;
    return d;
}

// ---------------------------------------------------------------------------------------------------------------------
// Minimal HDF5 emitter.
class h5_writer {
public:
    using scalar = std::variant<double, int64_t, std::string>;

    void set(const std::string& path, double v) { leaf(path).data = dataset{{}, 'f', bytes(&v, 8), 8}; }
    void set(const std::string& path, int64_t v) { leaf(path).data = dataset{{}, 'i', bytes(&v, 8), 8}; }
    void set(const std::string& path, int v) { int32_t w = v; leaf(path).data = dataset{{}, 'i', bytes(&w, 4), 4}; }
    void set(const std::string& path, bool v) { set(path, int(v)); }  // alps::hdf5 stores bool as an integer
    void set(const std::string& path, const std::string& v) {
        std::string raw = v;
        raw.push_back('\0');  // fixed length, null terminated: the terminator is part of the size
        leaf(path).data = dataset{{}, 's', raw, raw.size()};
    }
    void set(const std::string& path, const std::vector<double>& v, std::vector<uint64_t> shape = {}) {
        if (shape.empty()) shape = {v.size()};
        uint64_t n = 1;
        for (auto d : shape) n *= d;
        if (n != v.size()) throw std::logic_error("h5_writer: shape does not match the data of " + path);
        leaf(path).data = dataset{shape, 'f', bytes(v.data(), 8 * v.size()), 8};
    }
    void require_group(const std::string& path) { walk(path, true); }

    size_t save(const std::string& fname) {
        blob_.assign(96, '\0');  // the superblock goes in last
        const auto [oh, bt, hp] = write_group(root_);
        pad8();
        std::string sb("\x89HDF\r\n\x1a\n", 8);
        const unsigned char ver[8] = {0, 0, 0, 0, 0, 8, 8, 0};
        sb.append(reinterpret_cast<const char*>(ver), 8);
        put16(sb, leaf_k); put16(sb, internal_k); put32(sb, 0);
        put64(sb, 0); put64(sb, undef); put64(sb, blob_.size()); put64(sb, undef);
        put64(sb, 0); put64(sb, oh); put32(sb, 1); put32(sb, 0); put64(sb, bt); put64(sb, hp);
        std::memcpy(blob_.data(), sb.data(), 96);
        FILE* f = std::fopen(fname.c_str(), "wb");
        if (!f) throw std::runtime_error("h5_writer: cannot open " + fname);
        const size_t w = std::fwrite(blob_.data(), 1, blob_.size(), f);
        std::fclose(f);
        if (w != blob_.size()) throw std::runtime_error("h5_writer: short write to " + fname);
        return blob_.size();
    }

private:
    static constexpr uint64_t undef = ~uint64_t(0);
    static constexpr int leaf_k = 4, internal_k = 16;  // libhdf5 defaults
    struct dataset { std::vector<uint64_t> shape; char kind = 0; std::string raw; size_t itemsize = 0; };
    struct node { std::map<std::string, std::unique_ptr<node>> children; dataset data; bool is_group = true; };
    node root_;
    std::string blob_;

    static std::string bytes(const void* p, size_t n) { return std::string(reinterpret_cast<const char*>(p), n); }
    static void put16(std::string& s, uint16_t v) { s.append(reinterpret_cast<const char*>(&v), 2); }
    static void put32(std::string& s, uint32_t v) { s.append(reinterpret_cast<const char*>(&v), 4); }
    static void put64(std::string& s, uint64_t v) { s.append(reinterpret_cast<const char*>(&v), 8); }
    static void pad8(std::string& s) { while (s.size() % 8) s.push_back('\0'); }
    void pad8() { pad8(blob_); }
    uint64_t alloc(const std::string& data) {
        pad8();
        const uint64_t at = blob_.size();
        blob_ += data;
        return at;
    }
    node& walk(const std::string& path, bool groups_only) {
        node* g = &root_;
        size_t i = 0;
        while (i < path.size()) {
            while (i < path.size() && path[i] == '/') ++i;
            size_t j = path.find('/', i);
            if (j == std::string::npos) j = path.size();
            if (j > i) {
                if (!g->is_group) throw std::logic_error("h5_writer: " + path.substr(0, i) + " is a dataset, not a group");
                auto& ch = g->children[path.substr(i, j - i)];
                if (!ch) ch = std::make_unique<node>();
                g = ch.get();
            }
            i = j;
        }
        (void)groups_only;
        return *g;
    }
    node& leaf(const std::string& path) {
        node& n = walk(path, false);
        if (&n == &root_) throw std::logic_error("h5_writer: empty dataset path");
        n.is_group = false;
        return n;
    }
    static std::string message(uint16_t type, std::string data, uint8_t flags = 0) {
        pad8(data);
        std::string m;
        put16(m, type); put16(m, uint16_t(data.size())); m.push_back(char(flags)); m.append(3, '\0');
        return m + data;
    }
    static std::string object_header(const std::vector<std::string>& msgs) {
        std::string body;
        for (auto& m : msgs) body += m;
        std::string h;
        h.push_back(1); h.push_back(0); put16(h, uint16_t(msgs.size())); put32(h, 1); put32(h, uint32_t(body.size())); h.append(4, '\0');
        return h + body;
    }
    uint64_t write_dataset(const dataset& d) {
        const uint64_t addr = d.raw.empty() ? undef : alloc(d.raw);
        std::string space;  // version 1 dataspace
        space.push_back(1); space.push_back(char(d.shape.size())); space.push_back(0); space.append(5, '\0');
        for (auto s : d.shape) put64(space, s);
        std::string type;
        if (d.kind == 'f') {  // IEEE double, little endian: byte for byte what libhdf5 writes
            const unsigned char t[8] = {0x11, 0x20, 0x3f, 0x00, 8, 0, 0, 0};
            type.append(reinterpret_cast<const char*>(t), 8);
            put16(type, 0); put16(type, 64); type.push_back(52); type.push_back(11); type.push_back(0); type.push_back(52); put32(type, 1023);
        } else if (d.kind == 'i') {
            const unsigned char t[4] = {0x10, 0x08, 0x00, 0x00};
            type.append(reinterpret_cast<const char*>(t), 4);
            put32(type, uint32_t(d.itemsize)); put16(type, 0); put16(type, uint16_t(8 * d.itemsize));
        } else {
            const unsigned char t[4] = {0x13, 0x00, 0x00, 0x00};
            type.append(reinterpret_cast<const char*>(t), 4);
            put32(type, uint32_t(d.itemsize));
        }
        std::string fill("\x02\x02\x02\x00", 4);  // fill value v2: late allocation, written if set, undefined
        std::string layout;
        layout.push_back(3); layout.push_back(1); put64(layout, addr); put64(layout, d.raw.size());
        return alloc(object_header({message(0x0001, space), message(0x0003, type, 1), message(0x0005, fill), message(0x0008, layout)}));
    }
    std::array<uint64_t, 3> write_group(const node& g) {
        struct entry { std::string name; uint64_t oh; uint32_t cache; uint64_t bt, hp; };
        std::vector<entry> entries;  // std::map iterates in byte order of the names: the order the B-tree wants
        if (g.children.size() > size_t(4 * leaf_k * internal_k)) throw std::logic_error("h5_writer: too many members in one group");
        for (auto& [name, ch] : g.children) {
            if (ch->is_group) {
                const auto [oh, bt, hp] = write_group(*ch);
                entries.push_back({name, oh, 1, bt, hp});
            } else {
                entries.push_back({name, write_dataset(ch->data), 0, 0, 0});
            }
        }
        std::string heap(8, '\0');  // offset 0 holds the empty string
        std::vector<uint64_t> offs;
        for (auto& e : entries) {
            offs.push_back(heap.size());
            heap += e.name;
            heap.push_back('\0');
            pad8(heap);
        }
        const uint64_t heap_data = alloc(heap);
        std::string hh("HEAP");
        hh.push_back(0); hh.append(3, '\0'); put64(hh, heap.size()); put64(hh, 1); put64(hh, heap_data);  // free-list head 1 = none
        const uint64_t heap_addr = alloc(hh);
        std::vector<uint64_t> keys{0}, kids;
        for (size_t c0 = 0; c0 < entries.size(); c0 += 2 * leaf_k) {
            const size_t c1 = std::min(entries.size(), c0 + 2 * leaf_k);
            std::string sn("SNOD");
            sn.push_back(1); sn.push_back(0); put16(sn, uint16_t(c1 - c0));
            for (size_t i = c0; i < c1; ++i) {
                put64(sn, offs[i]); put64(sn, entries[i].oh); put32(sn, entries[i].cache); put32(sn, 0);
                put64(sn, entries[i].cache ? entries[i].bt : 0); put64(sn, entries[i].cache ? entries[i].hp : 0);
            }
            sn.append(40 * (2 * leaf_k - (c1 - c0)), '\0');
            kids.push_back(alloc(sn));
            keys.push_back(offs[c1 - 1]);
        }
        std::string tree("TREE");
        tree.push_back(0); tree.push_back(0); put16(tree, uint16_t(kids.size())); put64(tree, undef); put64(tree, undef);
        put64(tree, keys[0]);
        for (size_t i = 0; i < kids.size(); ++i) { put64(tree, kids[i]); put64(tree, keys[i + 1]); }
        tree.resize(24 + (2 * internal_k + 1) * 8 + 2 * internal_k * 8, '\0');
        const uint64_t bt_addr = alloc(tree);
        std::string st;
        put64(st, bt_addr); put64(st, heap_addr);
        const uint64_t oh_addr = alloc(object_header({message(0x0011, st)}));
        return {oh_addr, bt_addr, heap_addr};
    }
};

// [measurement][chain] (what fkmc_chain_get_series returns) -> one series, CHAIN-MAJOR: chain 0's measurements in Monte Carlo time
// order, then chain 1's, ... -- the order in which the reference gathers its ranks (src/measures/energy.cpp:32-47), so that bins run
// along MC time inside a chain instead of averaging neighbouring independent chains.
inline std::vector<double> pool_chains(const std::vector<double>& series, size_t n_measured, size_t n_chains) {
    if (series.size() != n_measured * n_chains) throw std::logic_error("pool_chains: series is not [n_measured][n_chains]");
    std::vector<double> out(series.size());
    for (size_t m = 0; m < n_measured; ++m)
        for (size_t c = 0; c < n_chains; ++c) out[c * n_measured + m] = series[m * n_chains + c];
    return out;
}

// ---------------------------------------------------------------------------------------------------------------------
using param_map = std::map<std::string, h5_writer::scalar>;

struct saved_stats {
    std::vector<std::array<double, 5>> binning;  // n, mean, variance, stderr, tau_int per bin level
    bin_stats_t stats;                           // the level estimate_bin picks
};

// prog/data_save.hxx (save_all_data -> save_measurements + energy / specific heat statistics) for the observables of the
// weight-evaluation path.  Series = all chains concatenated chain after chain (pool_chains), as the reference concatenates the ranks; they are binned in
// reverse order (rbegin..rend), like the reference.  histories: optional [index][measurement] tables for /mc_data.
inline std::map<std::string, saved_stats> save_all_data(const std::string& fname, const param_map& params, const std::vector<double>& energies,
                                                        const std::vector<double>& d2energies, const std::vector<double>& c_energies,
                                                        double beta, double volume, int max_depth = -1,
                                                        const std::map<std::string, std::pair<std::vector<double>, std::array<uint64_t, 2>>>& histories = {}) {
    h5_writer w;
    w.require_group("/parameters");
    for (auto& [k, v] : params) {
        const std::string path = "/parameters/" + k;
        if (std::holds_alternative<double>(v)) w.set(path, std::get<double>(v));
        else if (std::holds_alternative<int64_t>(v)) w.set(path, std::get<int64_t>(v));
        else w.set(path, std::get<std::string>(v));
    }
    w.set("/mc_data/energies", energies);
    w.set("/mc_data/d2energies", d2energies);
    w.set("/mc_data/c_energies", c_energies);
    for (auto& [k, h] : histories) w.set("/mc_data/" + k, h.first, {h.second[0], h.second[1]});
    if (max_depth < 0) max_depth = max_bin_depth(energies.size());
    std::map<std::string, saved_stats> out;
    auto put = [&](const std::string& name, const std::vector<bin_stats_t>& rows) {
        const auto cor = binning::calc_cor_length(rows);
        saved_stats s;
        std::vector<double> flat;
        for (size_t i = 0; i < rows.size(); ++i) {
            s.binning.push_back({rows[i][0], rows[i][1], rows[i][2], rows[i][3], cor[i]});
            flat.insert(flat.end(), s.binning.back().begin(), s.binning.back().end());
        }
        s.stats = rows[estimate_bin(rows)];
        w.set("/binning/" + name, flat, {uint64_t(rows.size()), 5});
        w.set("/stats/" + name, std::vector<double>(s.stats.begin(), s.stats.end()));
        out[name] = s;
    };
    auto rev = [](std::vector<double> v) { std::reverse(v.begin(), v.end()); return v; };
    const std::vector<double> e = rev(energies), d2 = rev(d2energies), ce = rev(c_energies);
    std::vector<double> e2(e.size());
    for (size_t i = 0; i < e.size(); ++i) e2[i] = e[i] * e[i];
    put("energy", binning::accumulate_binning(e, max_depth));
    put("d2energy", binning::accumulate_binning(d2, max_depth));
    put("c_energy", binning::accumulate_binning(ce, max_depth));
    // cv = beta^2 (<E^2> - <d2E> - <E>^2) / N  (prog/data_save.hxx:158-199)
    put("cv", jackknife::accumulate_jackknife([&](const std::vector<double>& m) { return beta * beta * (m[1] - m[2] - m[0] * m[0]) / volume; },
                                              {e, e2, d2}, max_depth));
    w.save(fname);
    return out;
}

}  // namespace fk
