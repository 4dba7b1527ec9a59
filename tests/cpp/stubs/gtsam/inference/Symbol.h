#pragma once
#include "../base.h"
namespace gtsam {
inline Key symbol(unsigned char c, std::uint64_t j) { return (static_cast<Key>(c) << 56) | j; }
namespace symbol_shorthand {
inline Key X(std::uint64_t j) { return symbol('x', j); }
inline Key G(std::uint64_t j) { return symbol('g', j); }
}  // namespace symbol_shorthand
}  // namespace gtsam
