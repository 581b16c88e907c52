// oracle/boost_shim -- TEST INFRASTRUCTURE ONLY.
// A stand-in for the few boost::geometry / boost::numeric::ublas names the reference's KITTI evaluator
// (scripts/offline_eval/kitti_native_eval/evaluate_object_3d_offline.cpp:11-23,267-345) uses, so that the evaluator
// compiles here FROM ITS OWN SOURCE without boost (not installed, no network).  Only what that file needs:
// convex quadrilaterals built by append(poly, points[5]), intersection(), union_() and area().
// intersection: Sutherland-Hodgman clipping of one convex polygon by the other;  union_: boost would return the merged
// outline -- the evaluator only ever takes area(un.front()), so the shim returns a polygon CARRYING that area
// (a + b - inter; for disjoint inputs boost returns both polygons and front() is the first one: area a).
// Everything else of the evaluator (data cleaning, matching, thresholds, precision / recall / AP) is the reference's.
#pragma once
#include <cmath>
#include <vector>

#define BOOST_GEOMETRY_REGISTER_C_ARRAY_CS(cs_)

namespace boost { namespace geometry {
namespace cs { struct cartesian {}; }
namespace model {
namespace d2 { template <typename T> struct point_xy { T x_, y_; }; }
template <typename P> struct polygon {
    std::vector<P> pts;          // open ring (the closing point is dropped)
    double carried_area = -1.0;  // >= 0: area() returns this (results of union_)
};
}  // namespace model

template <typename P>
void append(model::polygon<P>& poly, const double (&points)[5][2]) {
    for (int i = 0; i < 4; ++i) poly.pts.push_back(P{points[i][0], points[i][1]});
}

template <typename P>
double area(const model::polygon<P>& poly) {
    if (poly.carried_area >= 0) return poly.carried_area;
    double s = 0;
    const size_t n = poly.pts.size();
    for (size_t i = 0; i < n; ++i) {
        const P &a = poly.pts[i], &b = poly.pts[(i + 1) % n];
        s += a.x_ * b.y_ - b.x_ * a.y_;
    }
    return std::fabs(s) * 0.5;
}

namespace detail {
template <typename P>
double signed_area2(const std::vector<P>& r) {
    double s = 0;
    for (size_t i = 0; i < r.size(); ++i) {
        const P &a = r[i], &b = r[(i + 1) % r.size()];
        s += a.x_ * b.y_ - b.x_ * a.y_;
    }
    return s;
}
// clip `subject` by the convex ring `clip` (any orientation)
template <typename P>
std::vector<P> clip_convex(std::vector<P> subject, const std::vector<P>& clip) {
    const double orient = signed_area2(clip) >= 0 ? 1.0 : -1.0;
    for (size_t i = 0; i < clip.size() && !subject.empty(); ++i) {
        const P &c0 = clip[i], &c1 = clip[(i + 1) % clip.size()];
        const double ex = c1.x_ - c0.x_, ey = c1.y_ - c0.y_;
        std::vector<P> out;
        for (size_t j = 0; j < subject.size(); ++j) {
            const P &p = subject[j], &q = subject[(j + 1) % subject.size()];
            const double dp = orient * (ex * (p.y_ - c0.y_) - ey * (p.x_ - c0.x_));
            const double dq = orient * (ex * (q.y_ - c0.y_) - ey * (q.x_ - c0.x_));
            if (dp >= 0) out.push_back(p);
            if ((dp >= 0) != (dq >= 0)) {
                const double t = dp / (dp - dq);
                out.push_back(P{p.x_ + t * (q.x_ - p.x_), p.y_ + t * (q.y_ - p.y_)});
            }
        }
        subject.swap(out);
    }
    return subject;
}
}  // namespace detail

template <typename P>
void intersection(const model::polygon<P>& a, const model::polygon<P>& b, std::vector<model::polygon<P> >& out) {
    model::polygon<P> r;
    r.pts = detail::clip_convex(a.pts, b.pts);
    if (r.pts.size() >= 3 && area(r) > 0) out.push_back(r);
}

template <typename P>
void union_(const model::polygon<P>& a, const model::polygon<P>& b, std::vector<model::polygon<P> >& out) {
    std::vector<model::polygon<P> > in;
    intersection(a, b, in);
    model::polygon<P> r;
    r.carried_area = in.empty() ? area(a) : area(a) + area(b) - area(in.front());
    out.push_back(r);
}
}}  // namespace boost::geometry
