// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// Container-semantics stand-in for boost::dynamic_bitset<> (Boost 1.83 is pinned by the reference,
// MODULE.bazel:86-101, but is not vendored and there is no network).  Covers exactly the members
// the reference sources use (tesseract.h:18, tesseract.cc:73-82,231-252,396,436; visualization.cc).
#ifndef ORACLE_SHIM_BOOST_DYNAMIC_BITSET_HPP
#define ORACLE_SHIM_BOOST_DYNAMIC_BITSET_HPP
#include <cstddef>
#include <cstdint>
#include <vector>

namespace boost {

template <typename Block = uint64_t>
class dynamic_bitset {
 public:
  typedef size_t size_type;
  static constexpr size_type npos = static_cast<size_type>(-1);

  class reference {
   public:
    reference(Block& w, Block m) : w_(w), m_(m) {}
    operator bool() const {
      return (w_ & m_) != 0;
    }
    bool operator~() const {
      return (w_ & m_) == 0;
    }
    reference& operator=(bool v) {
      if (v)
        w_ |= m_;
      else
        w_ &= ~m_;
      return *this;
    }
    reference& operator=(const reference& o) {
      return *this = static_cast<bool>(o);
    }

   private:
    Block& w_;
    Block m_;
  };

  dynamic_bitset() : n_(0) {}
  explicit dynamic_bitset(size_type num_bits, unsigned long value = 0)
      : n_(num_bits), w_((num_bits + 63) / 64, 0) {
    if (value && !w_.empty()) w_[0] = value;
    trim();
  }

  size_type size() const {
    return n_;
  }
  reference operator[](size_type i) {
    return reference(w_[i >> 6], Block{1} << (i & 63));
  }
  bool operator[](size_type i) const {
    return (w_[i >> 6] >> (i & 63)) & 1;
  }
  dynamic_bitset& operator|=(const dynamic_bitset& o) {
    for (size_t k = 0; k < w_.size(); ++k) w_[k] |= o.w_[k];
    return *this;
  }
  dynamic_bitset& operator&=(const dynamic_bitset& o) {
    for (size_t k = 0; k < w_.size(); ++k) w_[k] &= o.w_[k];
    return *this;
  }
  dynamic_bitset operator~() const {
    dynamic_bitset r(*this);
    for (auto& w : r.w_) w = ~w;
    r.trim();
    return r;
  }
  size_type find_first() const {
    return scan_from(0);
  }
  size_type find_next(size_type pos) const {
    if (pos == npos || pos + 1 >= n_) return npos;
    return scan_from(pos + 1);
  }
  bool operator==(const dynamic_bitset& o) const {
    return n_ == o.n_ && w_ == o.w_;
  }
  bool operator!=(const dynamic_bitset& o) const {
    return !(*this == o);
  }
  const std::vector<Block>& words() const {
    return w_;
  }

 private:
  void trim() {
    if (n_ & 63) w_.back() &= (Block{1} << (n_ & 63)) - 1;
  }
  size_type scan_from(size_type start) const {
    size_type k = start >> 6;
    if (k >= w_.size()) return npos;
    Block cur = w_[k] & (~Block{0} << (start & 63));
    while (true) {
      if (cur) return (k << 6) + (size_type)__builtin_ctzll(cur);
      if (++k >= w_.size()) return npos;
      cur = w_[k];
    }
  }
  size_type n_;
  std::vector<Block> w_;
};

template <typename Block>
inline size_t hash_value(const dynamic_bitset<Block>& bs) {
  size_t h = bs.size();
  for (Block w : bs.words()) h ^= (size_t)w + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  return h;
}

}  // namespace boost
#endif
