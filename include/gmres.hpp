// gmres.hpp -- shape-compatible stand-in for the reference's `class Gmres` (include/gmres.hpp:8-129).
//
// In the reference, Gmres owns the Krylov workspace and runs the warm-started GMRES(k_max) on the host,
// calling back into the derived class through the virtual Ax_func.  Here the whole solve (Arnoldi with
// modified Gram-Schmidt, 2x2 Householder triangularisation, back substitution; gmres.hpp:28-112) is fused
// into the sm_100a control-update kernel and Ax_func is inlined device code, so this base class only keeps
// the constructor signature and the three parameters for source compatibility.  There is deliberately no
// host solver behind it: this framework has no CPU path.
#pragma once
#include <stdint.h>

class Gmres {
 protected:
  Gmres(const uint16_t len, const uint16_t k_max, const double tol) : len(len), k_max(k_max), tol(tol) {}
  ~Gmres() {}

  const uint16_t len;    // dim_u * dv
  const uint16_t k_max;  // Krylov dimension (Model::k_max)
  const double tol;      // convergence threshold (Model::tol)

 private:
  Gmres(const Gmres&);
  Gmres& operator=(const Gmres&);
};
