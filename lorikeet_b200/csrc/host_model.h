// host_model.h — host-side PairHMM model tables (see host_model.cpp).
#pragma once
#include <vector>

namespace lk {
struct ModelTables {
    std::vector<double> qerr_d;  // 10^(-q/10), q = 0..255
    std::vector<float> qerr_f;
    std::vector<double> m2m;     // triangular match->match, index max*(max+1)/2 + min, 0..255
};
void build_model_tables(ModelTables &t);
}  // namespace lk
