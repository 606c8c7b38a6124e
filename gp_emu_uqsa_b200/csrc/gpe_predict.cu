// K4/K5 prediction + implausibility (filled in below).
#include "gpe_handle.h"

using namespace gpe;

void gpe_handle::free_fit() {
    auto fr = [](double*& p) { if (p) cudaFree(p); p = nullptr; };
    fr(fLi); fr(fE); fr(fK); fr(fbeta); fr(fwinv); fr(fXs);
    fr(pC); fr(pPart); fr(pAux); fr(pX); fr(pH); fr(pMean); fr(pVar);
    pchunk = 0;
    fitted = false;
}

extern "C" {
int gpe_cross_cov(gpe_handle* h, const double*, double, int, const double*, int, double*) { return h ? h->fail_msg("not implemented") : -2; }
int gpe_fit_state(gpe_handle* h, const double*, double, double, int, const double*, double*, double*, int*) { return h ? h->fail_msg("not implemented") : -2; }
int gpe_predict(gpe_handle* h, const double*, const double*, long long, double*, double*) { return h ? h->fail_msg("not implemented") : -2; }
int gpe_predict_grid(gpe_handle* h, const int*, const double*, const double*, long long, long long, double*, double*) { return h ? h->fail_msg("not implemented") : -2; }
int gpe_predict_fullcov(gpe_handle* h, const double*, const double*, int, const double*, double*, double*) { return h ? h->fail_msg("not implemented") : -2; }
int gpe_implausibility(gpe_handle* h, const double*, const double*, int, long long, const double*, const double*, double, int, long long, double*, unsigned char*, unsigned long long*, double*, unsigned long long*) { return h ? h->fail_msg("not implemented") : -2; }
}
