// ORACLE (test infrastructure only) -- lattices and f-configuration helpers.
// Follows src/lattice/hypercubic.cpp and src/configuration.cpp of the reference.
#include <cmath>
#include <stdexcept>

#include "oracle.hpp"

namespace orc {

// src/lattice/hypercubic.cpp:31-39 -- last coordinate runs fastest
std::vector<int> lattice::index_to_pos(int index) const {
    std::vector<int> out(ndim);
    for (int i = ndim - 1; i >= 0; i--) {
        out[i] = index % L;
        index /= L;
    }
    return out;
}

// src/lattice/hypercubic.cpp:42-51
int lattice::pos_to_index(const std::vector<int>& pos) const {
    int out = 0, mult = 1;
    for (int i = ndim - 1; i >= 0; i--) {
        out += pos[i] * mult;
        mult *= L;
    }
    return out;
}

double lattice::hop(int i, int j) const {
    double v = 0;
    for (auto& e : rows[i])
        if (e.first == j) v += e.second;
    return v;
}

static void add_hopping(lattice& l, int i, int j, double v) {
    // src/lattice.cpp:9-17: SparseMatrix::insert -- an (i,j) pair may only be inserted once (Q6: L>=3)
    for (auto& e : l.rows[i])
        if (e.first == j) throw std::logic_error("oracle: duplicate hopping entry (L < 3?)");
    l.rows[i].push_back({j, v});
}

// src/lattice/hypercubic.cpp:116-131
static void fill_nearest_neighbors(lattice& l, double t) {
    for (int i = 0; i < l.N; ++i) {
        auto cur = l.index_to_pos(i);
        for (int n = 0; n < l.ndim; ++n) {
            auto pl = cur, pr = cur;
            pl[n] = (cur[n] > 0 ? cur[n] - 1 : l.L - 1);
            pr[n] = (cur[n] < l.L - 1 ? cur[n] + 1 : 0);
            add_hopping(l, i, l.pos_to_index(pl), -1.0 * t);
            add_hopping(l, i, l.pos_to_index(pr), -1.0 * t);
        }
    }
}

// src/lattice/hypercubic.cpp:137-155
static void fill_triangular(lattice& l, double t, double tp) {
    fill_nearest_neighbors(l, t);
    for (int i = 0; i < l.N; ++i) {
        auto cur = l.index_to_pos(i);
        auto pl = cur, pr = cur;
        for (int n = 0; n < 2; ++n) {
            pl[n] = (cur[n] > 0 ? cur[n] - 1 : l.L - 1);
            pr[n] = (cur[n] < l.L - 1 ? cur[n] + 1 : 0);
        }
        add_hopping(l, i, l.pos_to_index(pl), -1.0 * tp);
        add_hopping(l, i, l.pos_to_index(pr), -1.0 * tp);
    }
}

// src/lattice/hypercubic.cpp:160-203.  literal = true reproduces the per-linear-index toggle of
// the reference (non-symmetric for even L); literal = false is the intended brick wall.
static void fill_honeycomb(lattice& l, double t, bool literal) {
    const int X = 1, Y = 0;
    if (l.L % 2 != 0) throw std::logic_error("Need even size");
    bool subA = true;
    for (int i = 0; i < l.N; ++i) {
        auto cur = l.index_to_pos(i);
        auto pl = cur, pr = cur, pu = cur, pd = cur;
        pl[X] = (cur[X] > 0 ? cur[X] - 1 : l.L - 1);
        pr[X] = (cur[X] < l.L - 1 ? cur[X] + 1 : 0);
        pd[Y] = (cur[Y] > 0 ? cur[Y] - 1 : l.L - 1);
        pu[Y] = (cur[Y] < l.L - 1 ? cur[Y] + 1 : 0);
        add_hopping(l, i, l.pos_to_index(pl), -1.0 * t);
        add_hopping(l, i, l.pos_to_index(pr), -1.0 * t);
        bool a = literal ? subA : ((cur[X] + cur[Y]) % 2 == 0);
        if (a)
            add_hopping(l, i, l.pos_to_index(pu), -1.0 * t);
        else
            add_hopping(l, i, l.pos_to_index(pd), -1.0 * t);
        subA = !subA;
    }
}

lattice make_lattice(int kind, int L, double t, double tp) {
    lattice l;
    l.kind = kind;
    l.L = L;
    switch (kind) {
        case CUBIC1D: l.ndim = 1; break;
        case CUBIC2D: case TRIANGULAR: case HONEYCOMB: case HONEYCOMB_REF: case HONEYCOMB_REF_LOWER: l.ndim = 2; break;
        case CUBIC3D: l.ndim = 3; break;
        default: throw std::logic_error("oracle: unknown lattice kind");
    }
    if (L < 3) throw std::logic_error("oracle: L >= 3 required");
    l.N = 1;
    for (int d = 0; d < l.ndim; ++d) l.N *= L;
    l.rows.assign(l.N, {});
    switch (kind) {
        case CUBIC1D: case CUBIC2D: case CUBIC3D: fill_nearest_neighbors(l, t); break;
        case TRIANGULAR: fill_triangular(l, t, tp); break;
        case HONEYCOMB: fill_honeycomb(l, t, false); break;
        case HONEYCOMB_REF: fill_honeycomb(l, t, true); break;
        case HONEYCOMB_REF_LOWER: {
            fill_honeycomb(l, t, true);
            // keep the lower triangle (r >= c) of the literal matrix and mirror it
            lattice m = l;
            m.rows.assign(l.N, {});
            for (int r = 0; r < l.N; ++r)
                for (auto& e : l.rows[r])
                    if (r >= e.first) {
                        m.rows[r].push_back(e);
                        if (r != e.first) m.rows[e.first].push_back({r, e.second});
                    }
            l = m;
            break;
        }
    }
    return l;
}

// src/configuration.cpp:47-56
void randomize_f(random_generator& rnd, int V, size_t nf, std::vector<int>& f) {
    std::uniform_int_distribution<> distr(0, V - 1);
    if (!nf) nf = distr(rnd);
    f.assign(V, 0);
    for (size_t i = 0; i < nf; ++i) {
        size_t ind = distr(rnd);
        while (f[ind] == 1) ind = distr(rnd);
        f[ind] = 1;
    }
}

// src/configuration.cpp:59-77 (1-D only; 0 otherwise)
double calc_ff_energy(int ndim, const std::vector<int>& f, const std::vector<double>& W) {
    if (ndim != 1) return 0;
    double e = 0;
    int V = (int)f.size();
    for (int i = 0; i < V; ++i) {
        if (!f[i]) continue;
        for (int l = 0; l < (int)W.size(); ++l) {
            int left = (i - l + V) % V;
            int right = (i + l) % V;
            double el = W[l] * f[left];
            double er = W[l] * f[right];
            e += el;
            e += er * (l > 0);
        }
    }
    return e;
}

// src/configuration.cpp:79-91 followed by the sparse->dense copy at :212.  Column-major.
void dense_hamiltonian(const lattice& lat, const std::vector<int>& f, double U, double mu_c, std::vector<double>& H) {
    int n = lat.N;
    H.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) {
        for (auto& e : lat.rows[i]) H[(size_t)e.first * n + i] += e.second;  // H(i, j) at col j, row i
        H[(size_t)i * n + i] += -mu_c + U * f[i];
    }
}

}  // namespace orc
