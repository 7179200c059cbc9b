// sin and cos of an fp32 angle with |x| <= 8 rad (the range the FAST validity kernel analyses; larger joint values are
// sent to the exact kernel).  Three-term Cody-Waite reduction by pi/2 and the classic minimax polynomials on
// [-pi/4, pi/4]; no large-argument path, which is what keeps it small.  Error: checked exhaustively over every fp32
// value in [-8, 8] by tools/sincos_check.cu (result recorded in DESIGN.md); the FAST error analysis budgets 2 ulp(1.0).
#pragma once

namespace b200df
{
__device__ __forceinline__ void sincos_small(float x, float* s, float* c)
{
  const float kf = rintf(x * 0.636619772367581343f);  // x * 2/pi
  const int k = (int)kf;
  float r = fmaf(kf, -1.57079601287841796875f, x);      // pi/2 = hi + mid + lo
  r = fmaf(kf, -3.1391647326017846353352069854736328125e-7f, r);
  r = fmaf(kf, -5.390302529957764765544681e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  cp = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const bool swap = (k & 1) != 0;
  float ss = swap ? cp : sp, cc = swap ? sp : cp;
  if (k & 2)
    ss = -ss;
  if ((k + 1) & 2)
    cc = -cc;
  *s = ss;
  *c = cc;
}
}  // namespace b200df
