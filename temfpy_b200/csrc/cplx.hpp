// Minimal complex128 arithmetic for the complex-Slater kernels (usable in device code and in the CPU simulator).
#pragma once
#include "cta.hpp"

namespace tmf {
struct cplx {
  double x, y;
};
TMF_HD cplx cmake(double x, double y = 0.0) { cplx r; r.x = x; r.y = y; return r; }
TMF_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
TMF_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
TMF_HD cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
TMF_HD cplx cmulc(cplx a, cplx b) { return cmake(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }   // conj(a) * b
TMF_HD cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
TMF_HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
TMF_HD cplx cneg(cplx a) { return cmake(-a.x, -a.y); }
TMF_HD double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
TMF_HD cplx cinv(cplx a) {
  const double d = a.x * a.x + a.y * a.y;
  return (d > 0.0) ? cmake(a.x / d, -a.y / d) : cmake(0.0, 0.0);
}
TMF_HD cplx cdiv(cplx a, cplx b) { return cmul(a, cinv(b)); }
}  // namespace tmf
