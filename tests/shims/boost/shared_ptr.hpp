// TEST-ONLY stand-in for boost::shared_ptr (Boost is absent from this image): a DISTINCT template with the same surface,
// not convertible to or from std::shared_ptr, so a std/boost mix-up in the drop-in headers fails to compile exactly as it
// would against PCL <= 1.10, whose PointCloud<T>::Ptr is boost::shared_ptr. Not shipped.
#pragma once
#include <memory>
namespace boost {
template <typename T>
class shared_ptr {
 public:
  shared_ptr() {}
  template <typename U> explicit shared_ptr(U* p) : p_(p) {}
  template <typename U> shared_ptr(const shared_ptr<U>& o) : p_(o.std_()) {}
  T* get() const { return p_.get(); }
  T& operator*() const { return *p_; }
  T* operator->() const { return p_.get(); }
  explicit operator bool() const { return static_cast<bool>(p_); }
  void reset() { p_.reset(); }
  template <typename U> void reset(U* p) { p_.reset(p); }
  void swap(shared_ptr& o) { p_.swap(o.p_); }
  long use_count() const { return p_.use_count(); }
  const std::shared_ptr<T>& std_() const { return p_; }   // shim plumbing only
 private:
  std::shared_ptr<T> p_;
};
template <typename T, typename U> bool operator==(const shared_ptr<T>& a, const shared_ptr<U>& b) { return a.get() == b.get(); }
template <typename T, typename U> bool operator!=(const shared_ptr<T>& a, const shared_ptr<U>& b) { return a.get() != b.get(); }
}  // namespace boost
