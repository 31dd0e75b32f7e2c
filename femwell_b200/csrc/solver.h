// Shift-invert Arnoldi (thick restart) + post-processing entry points.
#pragma once
#include <complex>

#include "handle.h"

namespace fw {

bool dense_eig(int n, const std::complex<double>* a, std::complex<double>* w, std::complex<double>* v);

// device result: lams (host) and X_dev [k][n] c128 on the device (unit 2-norm Ritz vectors)
// returns false when fewer than k pairs converged: st->nconv pairs are then in lams[0:nconv] / X_dev[0:nconv] and
// *why holds the message (the API turns this into FW_ERR_NOCONV after handing the partial result to the caller)
bool eigs_shift_invert(Handle& h, const fw_eigs_params& prm, std::complex<double>* lams, DevBuf<double>& X_dev,
                       fw_eigs_stats* st, std::string* why = nullptr);

// y = (alphaA*A + alphaB*B) x with Dirichlet rows/cols masked (identity rows give 0 here)
template <class V, class S>
void op_apply(Handle& h, const uint8_t* active, S alphaA, S alphaB, const S* x, S* y);

void postprocess_modes_dev(Handle& h, int k, const std::complex<double>* lams, const double* X_dev, double k0,
                           double omega, double mu0, double* E_dev, double* H_dev, std::complex<double>* powers);
void hfield_dev(Handle& h, const cplx* E_dev, cplx beta, double omega, double mu0, cplx* H_dev);
void hfield_multi_dev(Handle& h, int k, const cplx* E_dev, const cplx* betas, double omega, double mu0, cplx* H_dev);
void mass_solve_multi(Handle& h, int k, const cplx* B, cplx* X);
void pcg_solve_multi(Handle& h, int n_rows, const int* indptr, const int* indices, const double* vals, int64_t nnz,
                     int k, const cplx* B, cplx* X, const int* pair = nullptr);
std::complex<double> overlap_cross_dev(Handle& h, const fw_space& sj, const cplx* Ei, const cplx* Hi, const double* Ej_host,
                                       const double* Hj_host, int mode, int n_sel, const int* elements_dev);
void intensity_dev(Handle& h, const cplx* E, const cplx* Hf, double* out);
void poynting_dev(Handle& h, const cplx* E, const cplx* Hf, cplx* out);
void energy_density_dev(Handle& h, const cplx* E, cplx* out);
void error_estimator_dev(Handle& h, const cplx* E, int eps_kind, bool eps_complex, const void* eps_dev, int nf,
                         const int* t2f_dev, const int* f2t_dev, int nqf, const double* s, const double* w,
                         const double* vec, double* eta_dev);
void setup_generic_csr(Handle& h, int n, const int* Kp, const int* Ki, const void* Kx, const int* Mp, const int* Mi,
                       const void* Mx, bool is_complex, const double* coords);
void functional_dev(Handle& h, int kind, const cplx* f0, const cplx* f1, const cplx* f2, const cplx* f3, int aux_kind,
                    bool aux_complex, const void* aux_dev, int n_sel, const int* elements_dev, int option,
                    std::complex<double>* out3);

}  // namespace fw
