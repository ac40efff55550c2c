// pseudo_loop -- the reference's pseudoknot-table class surface (src/pseudo_loop.hh:13-56) on top of the
// device-resident tables: same constructor, same public methods, public TriangleMatrix P.  The fill is one bulk GPU
// sweep (started on first use, see s_energy_matrix.hh); getters have Matrix4D::get / TriangleMatrix::get semantics
// on a host mirror; backtrack() processes ONE interval like the reference, by running that node of the traceback
// on the GPU (ccj_traceback_step) and prepending the nodes it pushes to stack_interval in the same order.
#ifndef CCJ_B200_PSEUDO_LOOP_HH
#define CCJ_B200_PSEUDO_LOOP_HH
#include <memory>
#include <string>

#include "s_energy_matrix.hh"

class pseudo_loop {
public:
    pseudo_loop(std::string seq, s_energy_matrix *V, short *S, short *S1, vrna_param_t *params);   // src/pseudo_loop.hh:17
    ~pseudo_loop();

    void compute_energies(cand_pos_t, cand_pos_t) {}   // src/pseudo_loop.cc:69-132, bulk on the GPU

    void backtrack(minimum_fold *f, seq_interval *cur_interval);   // src/pseudo_loop.cc:861-2820, one node
    void set_stack_interval(seq_interval *stack_interval) { this->stack_interval = stack_interval; }
    seq_interval *get_stack_interval() { return stack_interval; }
    std::string get_structure() { return structure; }
    minimum_fold *get_minimum_fold() { return f; }

    energy_t get_WB(cand_pos_t i, cand_pos_t j) { return wbwp(T2_WB, i, j); }   // src/pseudo_loop.cc:647-661
    energy_t get_WP(cand_pos_t i, cand_pos_t j) { return wbwp(T2_WP, i, j); }
    energy_t get_energy(cand_pos_t i, cand_pos_t j) { return P.get(i, j); }

    // declared by the reference (src/pseudo_loop.hh:35-38); only the M variant is defined there (:663-679)
    energy_t get_PfromLdoubleprime(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PfromRdoubleprime(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PfromMdoubleprime(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PfromOdoubleprime(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);

    // src/pseudo_loop.cc:682-820
    energy_t get_PLiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PLmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PRiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PRmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PMiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PMmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_POiloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_POmloop(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);

    TriangleMatrix P;   // src/pseudo_loop.hh:56

    // extensions (the reference keeps these tables private): any of the 22 gap tables (enum ccj_table4) with
    // Matrix4D::get semantics, and WBP / WPP
    energy_t get_gap(int table, cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l);
    energy_t get_PK(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PK, i, j, k, l); }
    energy_t get_PL(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PL, i, j, k, l); }
    energy_t get_PR(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PR, i, j, k, l); }
    energy_t get_PM(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PM, i, j, k, l); }
    energy_t get_PO(cand_pos_t i, cand_pos_t j, cand_pos_t k, cand_pos_t l) { return get_gap(T_PO, i, j, k, l); }
    energy_t get_WBP(cand_pos_t i, cand_pos_t j) { return WBP.get(i, j); }
    energy_t get_WPP(cand_pos_t i, cand_pos_t j) { return WPP.get(i, j); }

private:
    void insert_node(int i, int j, int k, int l, char type);   // src/pseudo_loop.cc:2823-2846
    energy_t wbwp(int table, cand_pos_t i, cand_pos_t j);
    cand_pos_t n;
    std::string seq;
    s_energy_matrix *V;
    seq_interval *stack_interval;
    std::string structure;
    minimum_fold *f;
    vrna_param_t *params_;
    short *S_;
    short *S1_;
    TriangleMatrix WPP, WBP;
    std::shared_ptr<ccj_shell_fold> fold_;
};
#endif
