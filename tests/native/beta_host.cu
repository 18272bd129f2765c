// Test harness (not part of the product library): exposes the SHIPPED decision logic of the beta bisection
// (aspire_b200/csrc/beta_math.cuh) to the CPU tests, which feed it numpy-computed sums and compare the beta it
// converges to with the reference's determine_beta.  Built by tests/test_beta_math_host.py with nvcc (host only).
#include "../../aspire_b200/csrc/beta_math.cuh"

extern "C" {
void beta_host_init(asp_beta_state_t* st, double beta_prev, double target_eff, double tol, double n_total, double d_min,
                    double d_max) {
    asp::beta_init_state(st, beta_prev, target_eff, tol, n_total, d_min, d_max);
}
void beta_host_decide(asp_beta_state_t* st) { asp::beta_decide(st); }
void beta_host_decide3(asp_beta_state_t* st, const double* sums) { asp::beta_decide3(st, sums); }
int beta_host_candidates(void) { return asp::kBetaCand; }
// root-guided search: the node the pass evaluates, the Newton update, the node decision
void beta_host_node(const asp_beta_state_t* st, double root, double* a, double* b, int* depth) {
    asp::beta_node(st, root, *a, *b, *depth);
}
void beta_host_newton(const asp_beta_state_t* st, double* scratch, double s1, double s2, double d1, double d2) {
    asp::beta_newton_update(st, scratch, s1, s2, d1, d2);
}
int beta_host_node_decide(asp_beta_state_t* st, double* scratch) { return asp::beta_node_decide(st, scratch); }
int beta_host_scratch_words(void) { return asp::kBsWords; }
int beta_host_scratch_root(void) { return asp::kBsRoot; }
}
