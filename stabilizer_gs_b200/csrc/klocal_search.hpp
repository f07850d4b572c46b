// klocal_search.hpp — host side of the 1D k-local dynamic programme (cpp/klocal_sparse.cpp,
// StateMachine cpp/state.cpp:445-550) and of the periodic driver (py/klocal_periodic.py).
#pragma once
#include <memory>

#include "common.hpp"
#include "hamiltonian.hpp"

namespace sgs {

class KlocalMachine {
  public:
    ~KlocalMachine();
    // periodic_cmax < 0: finite chain (generate_Sright_all); >= 0: generate_Sright_periodic(H, cmax)
    static Status create(const Hamiltonian &H, const sgs_opts &opts, int periodic_cmax, std::unique_ptr<KlocalMachine> &out,
                         int scale = 1);   // scale: capacity multiplier of the device batch buffers
    int m() const;
    int nstates() const;
    int sright_count(int m) const;
    Status evolve();
    Status ground_state(sgs_result *out);
    struct Impl;
    Impl *impl() { return impl_; }

  private:
    KlocalMachine();
    Impl *impl_;
};

Status klocal_search(const Hamiltonian &H, const sgs_opts &opts, sgs_result *out);
Status klocal_periodic(const Hamiltonian &H, int cmax, int c, const sgs_opts &opts, sgs_periodic_result *out);

}  // namespace sgs
