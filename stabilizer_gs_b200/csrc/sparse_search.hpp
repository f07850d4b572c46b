// sparse_search.hpp — device-resident plan of the sparse search (upload once, search many times).
#pragma once
#include <memory>

#include "common.hpp"
#include "hamiltonian.hpp"

namespace sgs {

class SparsePlan {
  public:
    ~SparsePlan();
    static Status create(const Hamiltonian &H, const sgs_opts &opts, std::unique_ptr<SparsePlan> &out);
    Status run(sgs_result *out);

  private:
    SparsePlan();
    struct Impl;
    Impl *impl_;
};

}  // namespace sgs
