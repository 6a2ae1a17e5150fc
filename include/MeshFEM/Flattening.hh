// Symmetric index flattening (mirrors Flattening.hh:21-83): Voigt order xx,yy,zz,yz,xz,xy.
#ifndef MESHFEM_B200_FLATTENING_HH
#define MESHFEM_B200_FLATTENING_HH
#include <cstddef>
#include <utility>

constexpr size_t flatLen(size_t dim) { return (dim * (dim + 1)) / 2; }
constexpr size_t flattenIndices(size_t dim, size_t i, size_t j) {
    return (i == j) ? i
                    : ((i < j) ? (dim * (dim + 1) - j * (j - 1)) / 2 - (i + 1)
                               : (dim * (dim + 1) - i * (i - 1)) / 2 - (j + 1));
}
template <size_t _Dim> inline constexpr size_t flattenIndices(size_t i, size_t j) { return flattenIndices(_Dim, i, j); }
using IdxPair = std::pair<size_t, size_t>;
template <size_t _Dim> inline IdxPair unflattenIndex(size_t i);
template <> inline IdxPair unflattenIndex<2>(size_t i) { return (i < 2) ? IdxPair{i, i} : IdxPair{0, 1}; }
template <> inline IdxPair unflattenIndex<3>(size_t i) {
    return (i < 3) ? IdxPair{i, i} : ((i == 3) ? IdxPair{1, 2} : ((i == 4) ? IdxPair{0, 2} : IdxPair{0, 1}));
}
#endif
