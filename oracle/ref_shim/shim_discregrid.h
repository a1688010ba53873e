// Stand-in for the parts of Discregrid (Q-Minh fork, absent here) that the reference's collision
// code uses.  TEST INFRASTRUCTURE.  KD-tree: median split on the longest axis down to single
// entities, one bounding sphere per node, breadth-first traversal gated by a predicate — the
// documented behaviour of Discregrid's KDTree; the exact tree shape upstream builds is unpinned
//  and only affects culling, not the contact set of the graded scenes.
#pragma once

#include "shim_eigen.h"

#include <array>
#include <functional>
#include <numeric>
#include <queue>
#include <vector>

namespace Discregrid {

class BoundingSphere
{
  public:
    BoundingSphere() : x_(0., 0., 0.), r_(0.) {}
    BoundingSphere(Eigen::Vector3d const& x, double r) : x_(x), r_(r) {}
    // enclosing ball of a point set (Ritter-style; upstream computes a minimal ball — any
    // enclosing ball is conservative for culling)
    explicit BoundingSphere(std::vector<Eigen::Vector3d> const& pts) : x_(0., 0., 0.), r_(0.)
    {
        if (pts.empty())
            return;
        Eigen::AlignedBox3d box;
        for (auto const& p : pts)
            box.extend(p);
        x_ = box.center();
        double r2 = 0.;
        for (auto const& p : pts)
            r2 = std::max(r2, (p - x_).squaredNorm());
        r_ = std::sqrt(r2) + 1e-10;
    }
    Eigen::Vector3d const& x() const { return x_; }
    Eigen::Vector3d& x() { return x_; }
    double r() const { return r_; }
    double& r() { return r_; }

  private:
    Eigen::Vector3d x_;
    double r_;
};

template <typename HullType>
class KDTree
{
  public:
    using TraversalPredicate = std::function<bool(unsigned int node_index, unsigned int depth)>;
    using TraversalCallback  = std::function<void(unsigned int node_index, unsigned int depth)>;

    struct Node
    {
        Node(unsigned int b_, unsigned int n_) : children({{-1, -1}}), begin(b_), n(n_) {}
        Node() = default;
        bool isLeaf() const { return children[0] < 0 && children[1] < 0; }
        std::array<int, 2> children{{-1, -1}};
        unsigned int begin = 0, n = 0;
    };

    explicit KDTree(std::size_t n) : m_lst(n) {}
    virtual ~KDTree() = default;

    Node const& node(unsigned int i) const { return m_nodes[i]; }
    HullType const& hull(unsigned int i) const { return m_hulls[i]; }
    unsigned int entity(unsigned int i) const { return m_lst[i]; }

    void construct()
    {
        m_nodes.clear();
        m_hulls.clear();
        if (m_lst.empty())
            return;
        std::iota(m_lst.begin(), m_lst.end(), 0u);
        Eigen::AlignedBox3d box;
        for (unsigned int i = 0; i < m_lst.size(); ++i)
            box.extend(entityPosition(i));
        add_node(0u, static_cast<unsigned int>(m_lst.size()));
        construct(0u, box, 0u, static_cast<unsigned int>(m_lst.size()));
    }

    void update()
    {
        for (unsigned int i = 0; i < m_nodes.size(); ++i)
            computeHull(m_nodes[i].begin, m_nodes[i].n, m_hulls[i]);
    }

    void traverseBreadthFirst(TraversalPredicate pred, TraversalCallback cb) const
    {
        if (m_nodes.empty())
            return;
        std::queue<std::pair<unsigned int, unsigned int>> pending;
        pending.push({0u, 0u});
        while (!pending.empty())
        {
            auto const [ni, depth] = pending.front();
            pending.pop();
            Node const& nd = m_nodes[ni];
            cb(ni, depth);
            bool const is_pred = pred(ni, depth);
            if (!nd.isLeaf() && is_pred)
            {
                pending.push({static_cast<unsigned int>(nd.children[0]), depth + 1});
                pending.push({static_cast<unsigned int>(nd.children[1]), depth + 1});
            }
        }
    }

  protected:
    virtual Eigen::Vector3d entityPosition(unsigned int i) const                     = 0;
    virtual void computeHull(unsigned int b, unsigned int n, HullType& hull) const = 0;

    std::vector<unsigned int> m_lst;

  private:
    int add_node(unsigned int b, unsigned int n)
    {
        HullType h;
        computeHull(b, n, h);
        m_hulls.push_back(h);
        m_nodes.push_back(Node(b, n));
        return static_cast<int>(m_nodes.size()) - 1;
    }
    void construct(unsigned int node, Eigen::AlignedBox3d const& box, unsigned int b, unsigned int n)
    {
        if (n <= 1)
            return;
        Eigen::Vector3d const d = box.diagonal();
        int axis = 0;
        if (d(1) > d(axis)) axis = 1;
        if (d(2) > d(axis)) axis = 2;
        std::sort(m_lst.begin() + b, m_lst.begin() + b + n, [&](unsigned int a, unsigned int c) {
            return entityPosition(a)(axis) < entityPosition(c)(axis);
        });
        unsigned int const hal = n / 2;
        Eigen::AlignedBox3d lbox, rbox;
        for (unsigned int i = b; i < b + hal; ++i)
            lbox.extend(entityPosition(m_lst[i]));
        for (unsigned int i = b + hal; i < b + n; ++i)
            rbox.extend(entityPosition(m_lst[i]));
        int const n0               = add_node(b, hal);
        m_nodes[node].children[0]  = n0;
        int const n1               = add_node(b + hal, n - hal);
        m_nodes[node].children[1]  = n1;
        construct(static_cast<unsigned int>(n0), lbox, b, hal);
        construct(static_cast<unsigned int>(n1), rbox, b + hal, n - hal);
    }
    std::vector<Node> m_nodes;
    std::vector<HullType> m_hulls;
};

// Grid SDF.  Discregrid's CubicLagrangeDiscreteGrid as published upstream (restated from its
// documented algorithm; source absent, PARITY UNPINNED): cells of 32 nodes (8 corners + 2 nodes on
// each of the 12 edges), per-cell node tables built by addFunction, serendipity cubic shape
// functions on [-1,1]^3, DBL_MAX outside the domain.  Written independently of oracle/xpbd_oracle.c
// (cell tables + unrolled shape functions here, index arithmetic + loops there) so that the two
// agreeing is a check of both.
class CubicLagrangeDiscreteGrid
{
  public:
    using ContinuousFunction = std::function<double(Eigen::Vector3d const&)>;
    CubicLagrangeDiscreteGrid(Eigen::AlignedBox3d const& domain, std::array<unsigned int, 3> const& resolution)
        : domain_(domain), resolution_(resolution)
    {
        for (int d = 0; d < 3; ++d)
        {
            cell_size_(d)     = resolution[d] ? domain.diagonal()(d) / static_cast<double>(resolution[d]) : 0.;
            inv_cell_size_(d) = cell_size_(d) != 0. ? 1. / cell_size_(d) : 0.;
        }
    }
    unsigned int addFunction(ContinuousFunction const& f)
    {
        auto const& n = resolution_;
        std::size_t const nv = (n[0] + 1) * (n[1] + 1) * (n[2] + 1);
        std::size_t const ne_x = n[0] * (n[1] + 1) * (n[2] + 1);
        std::size_t const ne_y = (n[0] + 1) * n[1] * (n[2] + 1);
        std::size_t const ne_z = (n[0] + 1) * (n[1] + 1) * n[2];
        nodes_.assign(nv + 2 * (ne_x + ne_y + ne_z), 0.);
        for (std::size_t l = 0; l < nodes_.size(); ++l)
            nodes_[l] = f(indexToNodePosition(static_cast<unsigned int>(l)));
        cells_.assign(static_cast<std::size_t>(n[0]) * n[1] * n[2], {});
        for (unsigned int l = 0; l < cells_.size(); ++l)
        {
            unsigned int const k = l / (n[1] * n[0]);
            unsigned int const t = l % (n[1] * n[0]);
            unsigned int const j = t / n[0];
            unsigned int const i = t % n[0];
            unsigned int const nx = n[0], ny = n[1], nz = n[2];
            auto& c = cells_[l];
            c[0] = (nx + 1) * (ny + 1) * k + (nx + 1) * j + i;
            c[1] = (nx + 1) * (ny + 1) * k + (nx + 1) * j + i + 1;
            c[2] = (nx + 1) * (ny + 1) * k + (nx + 1) * (j + 1) + i;
            c[3] = (nx + 1) * (ny + 1) * k + (nx + 1) * (j + 1) + i + 1;
            c[4] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * j + i;
            c[5] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * j + i + 1;
            c[6] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * (j + 1) + i;
            c[7] = (nx + 1) * (ny + 1) * (k + 1) + (nx + 1) * (j + 1) + i + 1;
            auto offset = static_cast<unsigned int>(nv);
            c[8]  = offset + 2 * (nx * (ny + 1) * k + nx * j + i);
            c[9]  = c[8] + 1;
            c[10] = offset + 2 * (nx * (ny + 1) * (k + 1) + nx * j + i);
            c[11] = c[10] + 1;
            c[12] = offset + 2 * (nx * (ny + 1) * k + nx * (j + 1) + i);
            c[13] = c[12] + 1;
            c[14] = offset + 2 * (nx * (ny + 1) * (k + 1) + nx * (j + 1) + i);
            c[15] = c[14] + 1;
            offset += 2 * static_cast<unsigned int>(ne_x);
            c[16] = offset + 2 * (ny * (nz + 1) * i + ny * k + j);
            c[17] = c[16] + 1;
            c[18] = offset + 2 * (ny * (nz + 1) * (i + 1) + ny * k + j);
            c[19] = c[18] + 1;
            c[20] = offset + 2 * (ny * (nz + 1) * i + ny * (k + 1) + j);
            c[21] = c[20] + 1;
            c[22] = offset + 2 * (ny * (nz + 1) * (i + 1) + ny * (k + 1) + j);
            c[23] = c[22] + 1;
            offset += 2 * static_cast<unsigned int>(ne_y);
            c[24] = offset + 2 * (nz * (nx + 1) * j + nz * i + k);
            c[25] = c[24] + 1;
            c[26] = offset + 2 * (nz * (nx + 1) * (j + 1) + nz * i + k);
            c[27] = c[26] + 1;
            c[28] = offset + 2 * (nz * (nx + 1) * j + nz * (i + 1) + k);
            c[29] = c[28] + 1;
            c[30] = offset + 2 * (nz * (nx + 1) * (j + 1) + nz * (i + 1) + k);
            c[31] = c[30] + 1;
        }
        return 0u;
    }
    double interpolate(unsigned int, Eigen::Vector3d const& x, Eigen::Vector3d* gradient = nullptr) const
    {
        if (cells_.empty() || !domain_.contains(x))
            return std::numeric_limits<double>::max();
        unsigned int mi[3];
        for (int d = 0; d < 3; ++d)
        {
            mi[d] = static_cast<unsigned int>((x(d) - domain_.min()(d)) * inv_cell_size_(d));
            if (mi[d] >= resolution_[d])
                mi[d] = resolution_[d] - 1;
        }
        unsigned int const i = resolution_[1] * resolution_[0] * mi[2] + resolution_[0] * mi[1] + mi[0];
        Eigen::Vector3d lo, hi, c0, xi;
        for (int d = 0; d < 3; ++d)
        {
            lo(d) = domain_.min()(d) + cell_size_(d) * static_cast<double>(mi[d]);
            hi(d) = lo(d) + cell_size_(d);
            c0(d) = 2.0 / (hi(d) - lo(d));
            xi(d) = c0(d) * x(d) - (hi(d) + lo(d)) / (hi(d) - lo(d));
        }
        double N[32], dN[32][3];
        shape(xi, N, gradient ? dN : nullptr);
        auto const& cell = cells_[i];
        double phi = 0.;
        for (int j = 0; j < 32; ++j)
            phi += N[j] * nodes_[cell[j]];
        if (gradient)
        {
            gradient->setZero();
            for (int j = 0; j < 32; ++j)
                for (int d = 0; d < 3; ++d)
                    (*gradient)(d) += dN[j][d] * nodes_[cell[j]];
            for (int d = 0; d < 3; ++d)
                (*gradient)(d) *= c0(d);
        }
        return phi;
    }
    Eigen::AlignedBox3d const& domain() const { return domain_; }
    std::array<unsigned int, 3> const& resolution() const { return resolution_; }
    std::vector<double> const& node_data() const { return nodes_; }
    Eigen::Vector3d indexToNodePosition(unsigned int l) const
    {
        auto const& n = resolution_;
        Eigen::Vector3d x;
        auto const nv   = (n[0] + 1) * (n[1] + 1) * (n[2] + 1);
        auto const ne_x = n[0] * (n[1] + 1) * (n[2] + 1);
        auto const ne_y = (n[0] + 1) * n[1] * (n[2] + 1);
        unsigned int ijk[3];
        if (l < nv)
        {
            ijk[2]      = l / ((n[1] + 1) * (n[0] + 1));
            auto temp   = l % ((n[1] + 1) * (n[0] + 1));
            ijk[1]      = temp / (n[0] + 1);
            ijk[0]      = temp % (n[0] + 1);
            for (int d = 0; d < 3; ++d)
                x(d) = domain_.min()(d) + cell_size_(d) * static_cast<double>(ijk[d]);
        }
        else if (l < nv + 2 * ne_x)
        {
            l -= nv;
            auto e_ind = l / 2;
            ijk[2]     = e_ind / ((n[1] + 1) * n[0]);
            auto temp  = e_ind % ((n[1] + 1) * n[0]);
            ijk[1]     = temp / n[0];
            ijk[0]     = temp % n[0];
            for (int d = 0; d < 3; ++d)
                x(d) = domain_.min()(d) + cell_size_(d) * static_cast<double>(ijk[d]);
            x(0) += (1.0 + static_cast<double>(l % 2)) / 3.0 * cell_size_(0);
        }
        else if (l < nv + 2 * (ne_x + ne_y))
        {
            l -= (nv + 2 * ne_x);
            auto e_ind = l / 2;
            ijk[0]     = e_ind / ((n[2] + 1) * n[1]);
            auto temp  = e_ind % ((n[2] + 1) * n[1]);
            ijk[2]     = temp / n[1];
            ijk[1]     = temp % n[1];
            for (int d = 0; d < 3; ++d)
                x(d) = domain_.min()(d) + cell_size_(d) * static_cast<double>(ijk[d]);
            x(1) += (1.0 + static_cast<double>(l % 2)) / 3.0 * cell_size_(1);
        }
        else
        {
            l -= (nv + 2 * (ne_x + ne_y));
            auto e_ind = l / 2;
            ijk[1]     = e_ind / ((n[0] + 1) * n[2]);
            auto temp  = e_ind % ((n[0] + 1) * n[2]);
            ijk[0]     = temp / n[2];
            ijk[2]     = temp % n[2];
            for (int d = 0; d < 3; ++d)
                x(d) = domain_.min()(d) + cell_size_(d) * static_cast<double>(ijk[d]);
            x(2) += (1.0 + static_cast<double>(l % 2)) / 3.0 * cell_size_(2);
        }
        return x;
    }

  private:
    static void shape(Eigen::Vector3d const& xi, double* res, double (*dN)[3])
    {
        double const x = xi(0), y = xi(1), z = xi(2);
        double const x2 = x * x, y2 = y * y, z2 = z * z;
        double const _1mx = 1.0 - x, _1my = 1.0 - y, _1mz = 1.0 - z;
        double const _1px = 1.0 + x, _1py = 1.0 + y, _1pz = 1.0 + z;
        double const _1m3x = 1.0 - 3.0 * x, _1m3y = 1.0 - 3.0 * y, _1m3z = 1.0 - 3.0 * z;
        double const _1p3x = 1.0 + 3.0 * x, _1p3y = 1.0 + 3.0 * y, _1p3z = 1.0 + 3.0 * z;
        double const _1mxt1my = _1mx * _1my, _1mxt1py = _1mx * _1py, _1pxt1my = _1px * _1my, _1pxt1py = _1px * _1py;
        double const _1mxt1mz = _1mx * _1mz, _1mxt1pz = _1mx * _1pz, _1pxt1mz = _1px * _1mz, _1pxt1pz = _1px * _1pz;
        double const _1myt1mz = _1my * _1mz, _1myt1pz = _1my * _1pz, _1pyt1mz = _1py * _1mz, _1pyt1pz = _1py * _1pz;
        double const _1mx2 = 1.0 - x2, _1my2 = 1.0 - y2, _1mz2 = 1.0 - z2;
        // corners
        double fac = 1.0 / 64.0 * (9.0 * (x2 + y2 + z2) - 19.0);
        res[0] = fac * _1mxt1my * _1mz;
        res[1] = fac * _1pxt1my * _1mz;
        res[2] = fac * _1mxt1py * _1mz;
        res[3] = fac * _1pxt1py * _1mz;
        res[4] = fac * _1mxt1my * _1pz;
        res[5] = fac * _1pxt1my * _1pz;
        res[6] = fac * _1mxt1py * _1pz;
        res[7] = fac * _1pxt1py * _1pz;
        // edges along x
        fac                 = 9.0 / 64.0 * _1mx2;
        double const f1m3x = fac * _1m3x, f1p3x = fac * _1p3x;
        res[8]  = f1m3x * _1myt1mz;
        res[9]  = f1p3x * _1myt1mz;
        res[10] = f1m3x * _1myt1pz;
        res[11] = f1p3x * _1myt1pz;
        res[12] = f1m3x * _1pyt1mz;
        res[13] = f1p3x * _1pyt1mz;
        res[14] = f1m3x * _1pyt1pz;
        res[15] = f1p3x * _1pyt1pz;
        // edges along y
        fac                 = 9.0 / 64.0 * _1my2;
        double const f1m3y = fac * _1m3y, f1p3y = fac * _1p3y;
        res[16] = f1m3y * _1mxt1mz;
        res[17] = f1p3y * _1mxt1mz;
        res[18] = f1m3y * _1pxt1mz;
        res[19] = f1p3y * _1pxt1mz;
        res[20] = f1m3y * _1mxt1pz;
        res[21] = f1p3y * _1mxt1pz;
        res[22] = f1m3y * _1pxt1pz;
        res[23] = f1p3y * _1pxt1pz;
        // edges along z
        fac                 = 9.0 / 64.0 * _1mz2;
        double const f1m3z = fac * _1m3z, f1p3z = fac * _1p3z;
        res[24] = f1m3z * _1mxt1my;
        res[25] = f1p3z * _1mxt1my;
        res[26] = f1m3z * _1mxt1py;
        res[27] = f1p3z * _1mxt1py;
        res[28] = f1m3z * _1pxt1my;
        res[29] = f1p3z * _1pxt1my;
        res[30] = f1m3z * _1pxt1py;
        res[31] = f1p3z * _1pxt1py;
        if (!dN)
            return;
        double const _9t3x2py2pz2m19 = 9.0 * (3.0 * x2 + y2 + z2) - 19.0;
        double const _9tx2p3y2pz2m19 = 9.0 * (x2 + 3.0 * y2 + z2) - 19.0;
        double const _9tx2py2p3z2m19 = 9.0 * (x2 + y2 + 3.0 * z2) - 19.0;
        double const _18x = 18.0 * x, _18y = 18.0 * y, _18z = 18.0 * z;
        double const _3m9x2 = 3.0 - 9.0 * x2, _3m9y2 = 3.0 - 9.0 * y2, _3m9z2 = 3.0 - 9.0 * z2;
        double const _2x = 2.0 * x, _2y = 2.0 * y, _2z = 2.0 * z;
        double const _18xm9t3x2py2pz2m19 = _18x - _9t3x2py2pz2m19, _18xp9t3x2py2pz2m19 = _18x + _9t3x2py2pz2m19;
        double const _18ym9tx2p3y2pz2m19 = _18y - _9tx2p3y2pz2m19, _18yp9tx2p3y2pz2m19 = _18y + _9tx2p3y2pz2m19;
        double const _18zm9tx2py2p3z2m19 = _18z - _9tx2py2p3z2m19, _18zp9tx2py2p3z2m19 = _18z + _9tx2py2p3z2m19;
        double const s = 1.0 / 64.0;
        dN[0][0] = _18xm9t3x2py2pz2m19 * _1myt1mz * s;
        dN[0][1] = _1mxt1mz * _18ym9tx2p3y2pz2m19 * s;
        dN[0][2] = _1mxt1my * _18zm9tx2py2p3z2m19 * s;
        dN[1][0] = _18xp9t3x2py2pz2m19 * _1myt1mz * s;
        dN[1][1] = _1pxt1mz * _18ym9tx2p3y2pz2m19 * s;
        dN[1][2] = _1pxt1my * _18zm9tx2py2p3z2m19 * s;
        dN[2][0] = _18xm9t3x2py2pz2m19 * _1pyt1mz * s;
        dN[2][1] = _1mxt1mz * _18yp9tx2p3y2pz2m19 * s;
        dN[2][2] = _1mxt1py * _18zm9tx2py2p3z2m19 * s;
        dN[3][0] = _18xp9t3x2py2pz2m19 * _1pyt1mz * s;
        dN[3][1] = _1pxt1mz * _18yp9tx2p3y2pz2m19 * s;
        dN[3][2] = _1pxt1py * _18zm9tx2py2p3z2m19 * s;
        dN[4][0] = _18xm9t3x2py2pz2m19 * _1myt1pz * s;
        dN[4][1] = _1mxt1pz * _18ym9tx2p3y2pz2m19 * s;
        dN[4][2] = _1mxt1my * _18zp9tx2py2p3z2m19 * s;
        dN[5][0] = _18xp9t3x2py2pz2m19 * _1myt1pz * s;
        dN[5][1] = _1pxt1pz * _18ym9tx2p3y2pz2m19 * s;
        dN[5][2] = _1pxt1my * _18zp9tx2py2p3z2m19 * s;
        dN[6][0] = _18xm9t3x2py2pz2m19 * _1pyt1pz * s;
        dN[6][1] = _1mxt1pz * _18yp9tx2p3y2pz2m19 * s;
        dN[6][2] = _1mxt1py * _18zp9tx2py2p3z2m19 * s;
        dN[7][0] = _18xp9t3x2py2pz2m19 * _1pyt1pz * s;
        dN[7][1] = _1pxt1pz * _18yp9tx2p3y2pz2m19 * s;
        dN[7][2] = _1pxt1py * _18zp9tx2py2p3z2m19 * s;
        double const _m3m9x2m2x = -_3m9x2 - _2x, _p3m9x2m2x = _3m9x2 - _2x;
        double const _1mx2t1m3x = _1mx2 * _1m3x, _1mx2t1p3x = _1mx2 * _1p3x;
        double const e = 9.0 / 64.0;
        dN[8][0]  = _m3m9x2m2x * _1myt1mz * e;  dN[8][1]  = -_1mx2t1m3x * _1mz * e; dN[8][2]  = -_1mx2t1m3x * _1my * e;
        dN[9][0]  = _p3m9x2m2x * _1myt1mz * e;  dN[9][1]  = -_1mx2t1p3x * _1mz * e; dN[9][2]  = -_1mx2t1p3x * _1my * e;
        dN[10][0] = _m3m9x2m2x * _1myt1pz * e;  dN[10][1] = -_1mx2t1m3x * _1pz * e; dN[10][2] = _1mx2t1m3x * _1my * e;
        dN[11][0] = _p3m9x2m2x * _1myt1pz * e;  dN[11][1] = -_1mx2t1p3x * _1pz * e; dN[11][2] = _1mx2t1p3x * _1my * e;
        dN[12][0] = _m3m9x2m2x * _1pyt1mz * e;  dN[12][1] = _1mx2t1m3x * _1mz * e;  dN[12][2] = -_1mx2t1m3x * _1py * e;
        dN[13][0] = _p3m9x2m2x * _1pyt1mz * e;  dN[13][1] = _1mx2t1p3x * _1mz * e;  dN[13][2] = -_1mx2t1p3x * _1py * e;
        dN[14][0] = _m3m9x2m2x * _1pyt1pz * e;  dN[14][1] = _1mx2t1m3x * _1pz * e;  dN[14][2] = _1mx2t1m3x * _1py * e;
        dN[15][0] = _p3m9x2m2x * _1pyt1pz * e;  dN[15][1] = _1mx2t1p3x * _1pz * e;  dN[15][2] = _1mx2t1p3x * _1py * e;
        double const _m3m9y2m2y = -_3m9y2 - _2y, _p3m9y2m2y = _3m9y2 - _2y;
        double const _1my2t1m3y = _1my2 * _1m3y, _1my2t1p3y = _1my2 * _1p3y;
        dN[16][0] = -_1my2t1m3y * _1mz * e; dN[16][1] = _m3m9y2m2y * _1mxt1mz * e; dN[16][2] = -_1my2t1m3y * _1mx * e;
        dN[17][0] = -_1my2t1p3y * _1mz * e; dN[17][1] = _p3m9y2m2y * _1mxt1mz * e; dN[17][2] = -_1my2t1p3y * _1mx * e;
        dN[18][0] = _1my2t1m3y * _1mz * e;  dN[18][1] = _m3m9y2m2y * _1pxt1mz * e; dN[18][2] = -_1my2t1m3y * _1px * e;
        dN[19][0] = _1my2t1p3y * _1mz * e;  dN[19][1] = _p3m9y2m2y * _1pxt1mz * e; dN[19][2] = -_1my2t1p3y * _1px * e;
        dN[20][0] = -_1my2t1m3y * _1pz * e; dN[20][1] = _m3m9y2m2y * _1mxt1pz * e; dN[20][2] = _1my2t1m3y * _1mx * e;
        dN[21][0] = -_1my2t1p3y * _1pz * e; dN[21][1] = _p3m9y2m2y * _1mxt1pz * e; dN[21][2] = _1my2t1p3y * _1mx * e;
        dN[22][0] = _1my2t1m3y * _1pz * e;  dN[22][1] = _m3m9y2m2y * _1pxt1pz * e; dN[22][2] = _1my2t1m3y * _1px * e;
        dN[23][0] = _1my2t1p3y * _1pz * e;  dN[23][1] = _p3m9y2m2y * _1pxt1pz * e; dN[23][2] = _1my2t1p3y * _1px * e;
        double const _m3m9z2m2z = -_3m9z2 - _2z, _p3m9z2m2z = _3m9z2 - _2z;
        double const _1mz2t1m3z = _1mz2 * _1m3z, _1mz2t1p3z = _1mz2 * _1p3z;
        dN[24][0] = -_1mz2t1m3z * _1my * e; dN[24][1] = -_1mz2t1m3z * _1mx * e; dN[24][2] = _m3m9z2m2z * _1mxt1my * e;
        dN[25][0] = -_1mz2t1p3z * _1my * e; dN[25][1] = -_1mz2t1p3z * _1mx * e; dN[25][2] = _p3m9z2m2z * _1mxt1my * e;
        dN[26][0] = -_1mz2t1m3z * _1py * e; dN[26][1] = _1mz2t1m3z * _1mx * e;  dN[26][2] = _m3m9z2m2z * _1mxt1py * e;
        dN[27][0] = -_1mz2t1p3z * _1py * e; dN[27][1] = _1mz2t1p3z * _1mx * e;  dN[27][2] = _p3m9z2m2z * _1mxt1py * e;
        dN[28][0] = _1mz2t1m3z * _1my * e;  dN[28][1] = -_1mz2t1m3z * _1px * e; dN[28][2] = _m3m9z2m2z * _1pxt1my * e;
        dN[29][0] = _1mz2t1p3z * _1my * e;  dN[29][1] = -_1mz2t1p3z * _1px * e; dN[29][2] = _p3m9z2m2z * _1pxt1my * e;
        dN[30][0] = _1mz2t1m3z * _1py * e;  dN[30][1] = _1mz2t1m3z * _1px * e;  dN[30][2] = _m3m9z2m2z * _1pxt1py * e;
        dN[31][0] = _1mz2t1p3z * _1py * e;  dN[31][1] = _1mz2t1p3z * _1px * e;  dN[31][2] = _p3m9z2m2z * _1pxt1py * e;
    }

    Eigen::AlignedBox3d domain_;
    std::array<unsigned int, 3> resolution_;
    Eigen::Vector3d cell_size_, inv_cell_size_;
    std::vector<double> nodes_;
    std::vector<std::array<unsigned int, 32>> cells_;
};
using CubicLagrangeGrid = CubicLagrangeDiscreteGrid;

class TriangleMesh
{
  public:
    TriangleMesh(std::vector<Eigen::Vector3d> const& v, std::vector<std::array<unsigned int, 3>> const& f)
        : vertices_(v), faces_(f)
    {
    }
    std::vector<Eigen::Vector3d> const& vertices() const { return vertices_; }
    std::vector<std::array<unsigned int, 3>> const& faces() const { return faces_; }

  private:
    std::vector<Eigen::Vector3d> vertices_;
    std::vector<std::array<unsigned int, 3>> faces_;
};

// MeshDistance: distance to the closest triangle (exhaustive here; upstream prunes the same search
// with a sphere hierarchy and caches results), signed by the angle-weighted pseudo-normal of the
// closest feature.  Closest-point computation by clamped barycentric minimisation (Eberly's
// region scheme), written independently of the Voronoi-region version in oracle/xpbd_oracle.c.
class MeshDistance
{
  public:
    explicit MeshDistance(TriangleMesh const& m) : mesh_(m)
    {
        auto const& X = m.vertices();
        auto const& F = m.faces();
        face_normals_.resize(F.size());
        vertex_normals_.assign(X.size(), Eigen::Vector3d(0., 0., 0.));
        for (std::size_t i = 0; i < F.size(); ++i)
        {
            Eigen::Vector3d const &x0 = X[F[i][0]], &x1 = X[F[i][1]], &x2 = X[F[i][2]];
            Eigen::Vector3d n = (x1 - x0).cross(x2 - x0);
            n.normalize();
            face_normals_[i] = n;
            Eigen::Vector3d e1 = x1 - x0, e2 = x2 - x1, e3 = x0 - x2;
            e1.normalize();
            e2.normalize();
            e3.normalize();
            double const alpha[3] = {std::acos(clamp1(-e1.dot(e3))), std::acos(clamp1(-e2.dot(e1))),
                                     std::acos(clamp1(-e3.dot(e2)))};
            for (int k = 0; k < 3; ++k)
                vertex_normals_[F[i][k]] = vertex_normals_[F[i][k]] + alpha[k] * n;
        }
    }
    double signedDistanceCached(Eigen::Vector3d const& x) const { return signedDistance(x); }
    double signedDistance(Eigen::Vector3d const& x) const
    {
        auto const& X = mesh_.vertices();
        auto const& F = mesh_.faces();
        double best = std::numeric_limits<double>::max();
        Eigen::Vector3d bp(0., 0., 0.);
        std::size_t bf = 0;
        double bs = 0., bt = 0.;
        for (std::size_t i = 0; i < F.size(); ++i)
        {
            double s, t;
            Eigen::Vector3d const q = closest(x, X[F[i][0]], X[F[i][1]], X[F[i][2]], s, t);
            double const d2         = (x - q).squaredNorm();
            if (d2 < best)
            {
                best = d2; bp = q; bf = i; bs = s; bt = t;
            }
        }
        if (F.empty())
            return best;
        // feature from the barycentric coordinates (1 - s - t, s, t) of the closest point
        double const b[3] = {1. - bs - bt, bs, bt};
        int zeros = 0, zi = -1, nzi = -1;
        for (int k = 0; k < 3; ++k)
            if (b[k] <= 0.) { ++zeros; zi = k; } else nzi = k;
        Eigen::Vector3d n = face_normals_[bf];
        if (zeros == 2)
            n = vertex_normals_[F[bf][nzi]];
        else if (zeros == 1)
        { // edge opposite to corner zi: between corners (zi+1)%3 and (zi+2)%3
            unsigned int const a = F[bf][(zi + 1) % 3], c = F[bf][(zi + 2) % 3];
            for (std::size_t o = 0; o < F.size(); ++o)
                if (o != bf)
                    for (int k = 0; k < 3; ++k)
                        if (F[o][k] == c && F[o][(k + 1) % 3] == a)
                        {
                            n = n + face_normals_[o];
                            o = F.size() - 1;
                            break;
                        }
        }
        double const dist = std::sqrt(best);
        return (x - bp).dot(n) < 0. ? -dist : dist;
    }

  private:
    static double clamp1(double c) { return c > 1. ? 1. : c < -1. ? -1. : c; }
    // minimise |a + s*e0 + t*e1 - p|^2 over s >= 0, t >= 0, s + t <= 1
    static Eigen::Vector3d closest(Eigen::Vector3d const& p, Eigen::Vector3d const& a, Eigen::Vector3d const& b,
                                   Eigen::Vector3d const& c, double& s, double& t)
    {
        Eigen::Vector3d const e0 = b - a, e1 = c - a, d = a - p;
        double const A = e0.dot(e0), B = e0.dot(e1), C = e1.dot(e1), D = e0.dot(d), E = e1.dot(d);
        double const det = A * C - B * B;
        s = B * E - C * D;
        t = B * D - A * E;
        auto const clamp01 = [](double v) { return v < 0. ? 0. : v > 1. ? 1. : v; };
        if (s + t <= det)
        {
            if (s < 0.)
            {
                if (t < 0.)
                { // region 4
                    if (D < 0.) { t = 0.; s = clamp01(-D / A); }
                    else { s = 0.; t = clamp01(-E / C); }
                }
                else { s = 0.; t = clamp01(-E / C); } // region 3
            }
            else if (t < 0.) { t = 0.; s = clamp01(-D / A); } // region 5
            else { s /= det; t /= det; }                       // region 0
        }
        else
        {
            if (s < 0.)
            { // region 2
                double const tmp0 = B + D, tmp1 = C + E;
                if (tmp1 > tmp0)
                {
                    double const numer = tmp1 - tmp0, denom = A - 2. * B + C;
                    s = clamp01(numer / denom);
                    t = 1. - s;
                }
                else { s = 0.; t = clamp01(-E / C); }
            }
            else if (t < 0.)
            { // region 6
                double const tmp0 = B + E, tmp1 = A + D;
                if (tmp1 > tmp0)
                {
                    double const numer = tmp1 - tmp0, denom = A - 2. * B + C;
                    t = clamp01(numer / denom);
                    s = 1. - t;
                }
                else { t = 0.; s = clamp01(-D / A); }
            }
            else
            { // region 1
                double const numer = (C + E) - (B + D), denom = A - 2. * B + C;
                s = clamp01(numer / denom);
                t = 1. - s;
            }
        }
        return a + s * e0 + t * e1;
    }
    TriangleMesh const& mesh_;
    std::vector<Eigen::Vector3d> face_normals_, vertex_normals_;
};

} // namespace Discregrid
