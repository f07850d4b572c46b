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

// ---- stepwise access to the DP states: what the Python classes State / StateMachinePeriodic / EnergyLevel of
// py/klocal_periodic.py need (include/sgs.h: sgs_klocal_states_*).  A KlocalStates is a host snapshot of a list of
// states of one site; it remembers which device state each entry was read from and the machine's history epoch, so
// that the history of a table entry (EnergyLevel.get_state) can be walked on the device while that is still valid.
struct KlocalStates;
void klocal_states_free(KlocalStates *s);
int klocal_states_count(const KlocalStates *s);
Status klocal_get_states(KlocalMachine &M, KlocalStates **out);
Status klocal_states_select(const KlocalStates *s, const int *idx, int count, KlocalStates **out);
Status klocal_states_append(KlocalStates *dst, const KlocalStates *src, int i);
Status klocal_state_info(const KlocalMachine &M, const KlocalStates *s, int i, int *m, int *rank, int *begin, int *end, int *sright);
Status klocal_state_rows(const KlocalMachine &M, const KlocalStates *s, int i, uint64_t *rows /* rank x W, absolute sites */);
Status klocal_state_key(const KlocalMachine &M, const KlocalStates *s, int i, uint64_t *key, int cap, int *nwords);
Status klocal_state_energies(const KlocalStates *s, int i, double *Es, int set);
Status klocal_states_shift(KlocalMachine &M, KlocalStates *s, int add, int clear_history);
Status klocal_set_states(KlocalMachine &M, const KlocalStates *s, int m);
Status klocal_state_history(KlocalMachine &M, const KlocalStates *s, int i, int entry, int *terms, int8_t *signs, int cap, int *count);
int klocal_sright_log(const KlocalMachine &M, int *pairs, int cap);
Status klocal_sright_group(KlocalMachine &M, int m, int i, int *rank, int *begin, int *end, uint64_t *rows, int8_t *signs);
Status klocal_state_valid_terms(const KlocalMachine &M, const KlocalStates *s, int i, int *terms, int cap, int *count);

Status klocal_search(const Hamiltonian &H, const sgs_opts &opts, sgs_result *out);
Status klocal_periodic(const Hamiltonian &H, int cmax, int c, const sgs_opts &opts, sgs_periodic_result *out);

}  // namespace sgs
