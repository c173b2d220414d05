// gwf_device.cuh -- device restatement of the stateless GWF helper functions:
//   hcond / vcond / condmean / logmean / staggered_thkfrac
//       src/Model/ModelUtilities/GwfConductanceUtils.f90:43-305, 375-393
//   sQuadraticSaturation(+Derivative), sQSaturation(+Derivative)
//       src/Utilities/SmoothingFunctions.f90:275-324, 364-406, 412-516
//   SsTerms / SyTerms / SsCapacity / SyCapacity
//       src/Model/ModelUtilities/GwfStorageUtils.f90:32-177
// Expression order is kept as in the reference so that (with -fmad=false) the
// results round identically.
#pragma once
#include "common.cuh"

namespace mf6 {
#ifdef __CUDACC__

__device__ __forceinline__ double sQuadraticSaturation(double top, double bot, double x, double eps) {
  double y;
  const double b = top - bot;
  if (b > 0.0) {
    double br;
    if (x < bot)
      br = 0.0;
    else if (x > top)
      br = 1.0;
    else
      br = (x - bot) / b;
    const double av = 1.0 / (1.0 - eps);
    const double bri = 1.0 - br;
    if (br < eps)
      y = av * 0.5 * (br * br) / eps;
    else if (br < (1.0 - eps))
      y = av * br + 0.5 * (1.0 - av);
    else if (br < 1.0)
      y = 1.0 - ((av * 0.5 * (bri * bri)) / eps);
    else
      y = 1.0;
  } else {
    y = (x < bot) ? 0.0 : 1.0;
  }
  return y;
}

__device__ __forceinline__ double sQuadraticSaturationDerivative(double top, double bot, double x,
                                                                 double eps) {
  const double b = top - bot;
  double br, y;
  if (x < bot)
    br = 0.0;
  else if (x > top)
    br = 1.0;
  else
    br = (x - bot) / b;
  const double av = 1.0 / (1.0 - eps);
  const double bri = 1.0 - br;
  if (br < eps)
    y = av * br / eps;
  else if (br < (1.0 - eps))
    y = av;
  else if (br < 1.0)
    y = av * bri / eps;
  else
    y = 0.0;
  return y / b;
}

__device__ __forceinline__ double sQSaturation(double top, double bot, double x) {
  const double w = x - bot, b = top - bot, s = w / b;
  const double cof1 = -2.0 / (b * b * b), cof2 = 3.0 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w * w) + cof2 * (w * w);
  return 1.0;
}

__device__ __forceinline__ double sQSaturationDerivative(double top, double bot, double x) {
  const double w = x - bot, b = top - bot, s = w / b;
  const double cof1 = -2.0 * 3.0 / (b * b * b), cof2 = 3.0 * 2.0 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w) + cof2 * w;
  return 0.0;
}

// the same two functions with explicit c1 / c2 (DRN passes -1, 2: gwf-drn.f90:563, 451-452)
__device__ __forceinline__ double sQSaturationC(double top, double bot, double x, double c1, double c2) {
  const double w = x - bot, b = top - bot, s = w / b;
  const double cof1 = c1 / (b * b * b), cof2 = c2 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w * w) + cof2 * (w * w);
  return 1.0;
}

__device__ __forceinline__ double sQSaturationDerivativeC(double top, double bot, double x, double c1,
                                                          double c2) {
  const double w = x - bot, b = top - bot, s = w / b;
  const double cof1 = c1 * 3.0 / (b * b * b), cof2 = c2 * 2.0 / (b * b);
  if (s < 0.0) return 0.0;
  if (s < 1.0) return cof1 * (w * w) + cof2 * w;
  return 0.0;
}

// get_drain_elevations (gwf-drn.f90:501-530): elev = b1, drndepth = the DDRN auxiliary value (0 = none)
__device__ __forceinline__ void drain_elevations(double drnelev, double drndepth, double &drntop,
                                                 double &drnbot) {
  if (drndepth != 0.0) {
    const double elev = drnelev + drndepth;
    drntop = fmax(elev, drnelev);
    drnbot = fmin(elev, drnelev);
  } else {
    drntop = drnelev;
    drnbot = drnelev;
  }
}

__device__ __forceinline__ double logmean(double d1, double d2) {
  const double drat = d2 / d1;
  if (drat <= 0.995 || drat >= 1.005) return (d2 - d1) / log(drat);
  return 0.5 * (d1 + d2);
}

__device__ __forceinline__ double condmean(double k1, double k2, double thick1, double thick2,
                                           double cl1, double cl2, double width, int iavgmeth) {
  const double t1 = k1 * thick1, t2 = k2 * thick2;
  double tmean, kmean, denom;
  switch (iavgmeth) {
    case 0:
      if (t1 * t2 > 0.0) return width * t1 * t2 / (t1 * cl2 + t2 * cl1);
      return 0.0;
    case 1:
      tmean = (t1 * t2 > 0.0) ? logmean(t1, t2) : 0.0;
      return tmean * width / (cl1 + cl2);
    case 2:
      kmean = (k1 * k2 > 0.0) ? logmean(k1, k2) : 0.0;
      return kmean * 0.5 * (thick1 + thick2) * width / (cl1 + cl2);
    case 3:
      denom = (k1 * cl2 + k2 * cl1);
      kmean = (denom > 0.0) ? k1 * k2 / denom : 0.0;
      return kmean * 0.5 * (thick1 + thick2) * width;
  }
  return 0.0;
}

__device__ __forceinline__ double staggered_thkfrac(double top, double bot, double sat, double topc,
                                                    double botc) {
  const double sill_top = fmin(top, topc), sill_bot = fmax(bot, botc);
  const double tp = bot + sat * (top - bot);
  return fmax(fmin(tp, sill_top) - sill_bot, 0.0);
}

__device__ __forceinline__ double hcond(int ibdn, int ibdm, int ictn, int ictm, int iupstream,
                                        int ihc, int icellavg, double condsat, double hn, double hm,
                                        double satn, double satm, double hkn, double hkm,
                                        double topn, double topm, double botn, double botm,
                                        double cln, double clm, double fawidth) {
  if (ibdn == 0 || ibdm == 0) return 0.0;
  if (ictn == 0 && ictm == 0) return condsat;
  if (iupstream == 1) {
    const double sat_up = (hn > hm) ? satn : satm;
    return sat_up * condsat;
  }
  double thksatn, thksatm;
  if (ihc == 2) {
    thksatn = staggered_thkfrac(topn, botn, satn, topm, botm);
    thksatm = staggered_thkfrac(topm, botm, satm, topn, botn);
  } else {
    thksatn = satn * (topn - botn);
    thksatm = satm * (topm - botm);
  }
  return condmean(hkn, hkm, thksatn, thksatm, cln, clm, fawidth, icellavg);
}

__device__ __forceinline__ double vcond(int ibdn, int ibdm, int ictn, int ictm, int ivarcv,
                                        int idewatcv, double condsat, double hn, double hm,
                                        double vkn, double vkm, double satn, double satm,
                                        double topn, double topm, double botn, double botm,
                                        double flowarea) {
  if (ibdn == 0 || ibdm == 0) return 0.0;
  if (ivarcv == 0) return condsat;
  if (ictn == 0 && ictm == 0) return condsat;
  if (hn >= topn && hm >= topm) return condsat;
  double satntmp = satn, satmtmp = satm;
  if (idewatcv == 0) {
    if (botn > botm)
      satmtmp = 1.0;
    else
      satntmp = 1.0;
  }
  const double bovk1 = satntmp * (topn - botn) * 0.5 / vkn;
  const double bovk2 = satmtmp * (topm - botm) * 0.5 / vkm;
  const double denom = bovk1 + bovk2;
  return (denom != 0.0) ? flowarea / denom : 0.0;
}

__device__ __forceinline__ double SsCapacity(int istor_coef, double top, double bot, double area,
                                             double ss) {
  const double thick = (istor_coef == 0) ? (top - bot) : 1.0;
  return ss * thick * area;
}

__device__ __forceinline__ void SsTerms(int iconvert, int iorig_ss, int iconf_ss, double top,
                                        double bot, double rho1, double rho1old, double snnew,
                                        double snold, double hnew, double hold, double &aterm,
                                        double &rhsterm, double &rate) {
  aterm = -rho1 * snnew;
  rhsterm = 0.0;
  if (iconvert != 0) {
    if (iorig_ss == 0) {
      if (iconf_ss == 0) {
        const double tthk = top - bot;
        const double zold = bot + 0.5 * tthk * snold;
        const double znew = bot + 0.5 * tthk * snnew;
        rhsterm = -rho1old * snold * (hold - zold) - rho1 * snnew * znew;
      } else {
        if (snold == 1.0) rhsterm = rhsterm - rho1old * (hold - top);
        if (snnew == 1.0)
          rhsterm = rhsterm - rho1 * top;
        else
          aterm = 0.0;
      }
    } else {
      rhsterm = -rho1old * snold * hold;
    }
  } else {
    rhsterm = -rho1old * snold * hold;
  }
  rate = aterm * hnew - rhsterm;
}

__device__ __forceinline__ void SyTerms(double top, double bot, double rho2, double rho2old,
                                        double snnew, double snold, double &aterm, double &rhsterm,
                                        double &rate) {
  aterm = 0.0;
  const double tthk = top - bot;
  if (snnew < 1.0) {
    if (snnew > 0.0) {
      aterm = -rho2;
      rhsterm = -rho2old * tthk * snold - rho2 * bot;
    } else {
      rhsterm = tthk * (0.0 - rho2old * snold);
    }
  } else {
    rhsterm = tthk * (rho2 * snnew - rho2old * snold);
  }
  rate = rho2old * tthk * snold - rho2 * tthk * snnew;
}

#endif
}  // namespace mf6
