// FRI arity-2 fold of one pair (third-party p3-fri fold_even_odd, reached from machine.prove):
//   out = (1/2 + t) a + (1/2 - t) b = (a + b)/2 + t (a - b),   t = beta / (2 x_i),
// x_i = w_{2^(logh+1)}^bitrev(i) the point of the pair's first element (the coset shift has been
// divided out by the reduced-opening step).  half_beta = beta / 2, ginv = 1 / w_{2^(logh+1)}.
#pragma once
#include "field.cuh"

namespace lsk {
LSK_HD Ext fri_fold_pair(const Ext& a, const Ext& b, const Ext& half_beta, uint32_t one_half, uint32_t ginv, uint32_t i,
                         unsigned logh) {
  Ext t = emul_base(half_beta, fpow(ginv, bitrev32(i, logh)));
  return eadd(emul_base(eadd(a, b), one_half), emul(t, esub(a, b)));
}
}  // namespace lsk
