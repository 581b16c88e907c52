// oracle/boost_shim -- TEST INFRASTRUCTURE ONLY: the dense row-major matrix<double>, .data() and prod() that the
// reference's KITTI evaluator uses to rotate box corners (evaluate_object_3d_offline.cpp:267-290).
#pragma once
#include <vector>
namespace boost { namespace numeric { namespace ublas {
template <typename T> class matrix {
  public:
    matrix(size_t r, size_t c) : r_(r), c_(c), d_(r * c) {}
    T& operator()(size_t i, size_t j) { return d_[i * c_ + j]; }
    const T& operator()(size_t i, size_t j) const { return d_[i * c_ + j]; }
    std::vector<T>& data() { return d_; }
    size_t size1() const { return r_; }
    size_t size2() const { return c_; }
  private:
    size_t r_, c_;
    std::vector<T> d_;
};
template <typename T> matrix<T> prod(const matrix<T>& a, const matrix<T>& b) {
    matrix<T> o(a.size1(), b.size2());
    for (size_t i = 0; i < a.size1(); ++i)
        for (size_t j = 0; j < b.size2(); ++j) {
            T s = 0;
            for (size_t k = 0; k < a.size2(); ++k) s += a(i, k) * b(k, j);
            o(i, j) = s;
        }
    return o;
}
}}}  // namespace boost::numeric::ublas
