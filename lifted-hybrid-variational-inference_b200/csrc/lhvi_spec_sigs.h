// Signatures with a template specialisation (no hidden discrete argument, K <= 3, T == 3).
//   full: X(NC, NG, NE)   two or more integrated arguments, F = log psi - log b; (0, 1, NE): a factor whose only
//                         integrated argument is a Gaussian-evidence class (C2F rounds), no gradient
//   pure: X(NC, NE)       unary split records, F = log psi
//   node: (NC, NG) = (1, 0) and (0, 1)
#pragma once
#define LHVI_SPEC_FULL(X) X(2, 0, 0) X(2, 0, 1) X(1, 1, 0) X(1, 1, 1) X(2, 1, 0) X(3, 0, 0) X(0, 1, 0) X(0, 1, 1)
#define LHVI_SPEC_PURE(X) X(0, 1) X(0, 2) X(1, 0) X(1, 1) X(1, 2)
