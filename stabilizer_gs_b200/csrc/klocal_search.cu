// klocal_search.cu — placeholder until the DP kernels land (replaced below in this round).
#include "klocal_search.hpp"
namespace sgs {
struct KlocalMachine::Impl {};
KlocalMachine::KlocalMachine() : impl_(new Impl()) {}
KlocalMachine::~KlocalMachine() { delete impl_; }
Status KlocalMachine::create(const Hamiltonian &, const sgs_opts &, int, std::unique_ptr<KlocalMachine> &) { return Status::fail(SGS_ERR_ARG, "k-local DP not built yet"); }
int KlocalMachine::m() const { return 0; }
int KlocalMachine::nstates() const { return 0; }
int KlocalMachine::sright_count(int) const { return 0; }
Status KlocalMachine::evolve() { return Status::fail(SGS_ERR_ARG, "k-local DP not built yet"); }
Status KlocalMachine::ground_state(sgs_result *) { return Status::fail(SGS_ERR_ARG, "k-local DP not built yet"); }
Status klocal_search(const Hamiltonian &, const sgs_opts &, sgs_result *) { return Status::fail(SGS_ERR_ARG, "k-local DP not built yet"); }
Status klocal_periodic(const Hamiltonian &, int, int, const sgs_opts &, sgs_periodic_result *) { return Status::fail(SGS_ERR_ARG, "k-local DP not built yet"); }
}
