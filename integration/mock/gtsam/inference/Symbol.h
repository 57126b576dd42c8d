// MOCK (see gtsam/base/Matrix.h)
#pragma once
#include <cstdint>
namespace gtsam { using Key = std::uint64_t;
namespace symbol_shorthand {
inline Key X(std::uint64_t j) { return (std::uint64_t('x') << 56) | j; }
inline Key V(std::uint64_t j) { return (std::uint64_t('v') << 56) | j; }
inline Key B(std::uint64_t j) { return (std::uint64_t('b') << 56) | j; }
} }
