// Host build of the reference kernels (TEST INFRASTRUCTURE, see cuda_shim.h).  build.py replaces the include line
// below by the text of /root/reference/kernels3.cu (with ONE line changed: `extern __shared__ double res[];` ->
// a function-local static array, which the shim's `__shared__` cannot express) and pipes the unit to g++;
// everything else is compiled exactly as it lies under /root/reference.
#include "cuda_shim.h"
#include REF_KERNELS

namespace {
// slot layout used by this repo: int32[14][n], field order of the reference struct (kernels3.cu:9-24)
frag view(int* base, int n) {
    frag f;
    f.pos = base + 0 * n; f.id_c = base + 1 * n; f.start_bp = base + 2 * n; f.len_bp = base + 3 * n;
    f.circ = base + 4 * n; f.id = base + 5 * n; f.prev = base + 6 * n; f.next = base + 7 * n;
    f.l_cont = base + 8 * n; f.l_cont_bp = base + 9 * n; f.ori = base + 10 * n; f.rep = base + 11 * n;
    f.activ = base + 12 * n; f.id_d = base + 13 * n;
    return f;
}
unsigned blocks(int n, int b) { return (unsigned)(n / b + 1); }      // the reference's grid: n_frags // size_block + 1
}

extern "C" {

// op: 0 flip, 1 swap_activity, 2 pop_out, 3..6 pop_in_1..4, 7 split, 8 paste, 9 simple_copy, 10 copy_struct
int emu_move(int op, int* dst, int* src, int* ids, int id_a, int id_b, int aux, int max_id, int n, int block) {
    frag fd = view(dst, n), fs = view(src, n);
    frag* d = &fd; frag* s = &fs;
    const unsigned g = blocks(n, block);
    switch (op) {
        case 0: emu_launch(g, 1, block, [&] { flip_frag(d, s, id_a, n); }); break;
        case 1: emu_launch(g, 1, block, [&] { swap_activity_frag(d, s, id_a, max_id, n); }); break;
        case 2: emu_launch(g, 1, block, [&] { pop_out_frag(d, s, ids, id_a, max_id, n); }); break;
        case 3: emu_launch(g, 1, block, [&] { pop_in_frag_1(d, s, id_a, id_b, max_id, aux, n); }); break;
        case 4: emu_launch(g, 1, block, [&] { pop_in_frag_2(d, s, id_a, id_b, max_id, aux, n); }); break;
        case 5: emu_launch(g, 1, block, [&] { pop_in_frag_3(d, s, id_a, id_b, max_id, aux, n); }); break;
        case 6: emu_launch(g, 1, block, [&] { pop_in_frag_4(d, s, id_a, id_b, max_id, aux, n); }); break;
        case 7: emu_launch(g, 1, block, [&] { split_contig(d, s, ids, id_a, aux, max_id, n); }); break;
        case 8: emu_launch(g, 1, block, [&] { paste_contigs(d, s, id_a, id_b, max_id, n); }); break;
        case 9: emu_launch(g, 1, block, [&] { simple_copy(d, s, n); }); break;
        case 10: emu_launch(g, 1, block, [&] { copy_struct(d, s, ids, n); }); break;
        default: return -1;
    }
    return 0;
}

int emu_fill_sub_index(int* src, int* sub_index, int contig_a, int contig_b, int l_cont_a, int n, int block) {
    frag fs = view(src, n);
    frag* s = &fs;
    const unsigned g = blocks(n, block);
    emu_launch(g, 1, block, [&] { fill_sub_index_fA(s, sub_index, contig_a, n); });
    if (contig_b != contig_a) emu_launch(g, 1, block, [&] { fill_sub_index_fB(s, sub_index, contig_b, l_cont_a, n); });
    return 0;
}

// evaluate_likelihood (kernels3.cu:2802-3222): per-pixel log-likelihood of a slot -> likelihood[max_id]
int emu_evaluate_likelihood(const float* obs, int* slot, int n, int* collector, int* dispatcher, int* id_sub, int* rep_id_sub,
                            float* len_sub, int* accu_sub, double* likelihood, const float* params8,
                            int max_id_up_diag, int max_id, int n_bins, int width, float nfpb, int grid, int block) {
    frag fs = view(slot, n);
    frag* s = &fs;
    param_simu P;
    P.kuhn = params8[0]; P.lm = params8[1]; P.c1 = params8[2]; P.slope = params8[3]; P.d = params8[4];
    P.d_max = params8[5]; P.fact = params8[6]; P.v_inter = params8[7];
    emu_launch((unsigned)grid, 1, block, [&] {
        evaluate_likelihood(obs, s, collector, (int2*)dispatcher, (int4*)id_sub, (int4*)rep_id_sub, (float3*)len_sub, (int3*)accu_sub,
                            likelihood, &P, max_id_up_diag, max_id, n_bins, width, nfpb);
    });
    return 0;
}

// sub_compute_likelihood (kernels3.cu:3259-3718): likelihood[0] += sum over the touched pixels of new - old
int emu_sub_compute_likelihood(const float* obs, int* slot, int n, int* sub_index, int* list_rep, int* list_uniq, int* collector,
                               int* dispatcher, int* id_sub, float* len_sub, int* accu_sub, double* likelihood, double* curr_likelihood,
                               const float* params8, int max_id_no_repeats, int lim_repeats_vs_uniq, int lim_intra_repeats, int max_id,
                               int n_frags_uniq, int n_repeats, int width, int n_bins, float nfpb, int grid, int block) {
    frag fs = view(slot, n);
    frag* s = &fs;
    param_simu P;
    P.kuhn = params8[0]; P.lm = params8[1]; P.c1 = params8[2]; P.slope = params8[3]; P.d = params8[4];
    P.d_max = params8[5]; P.fact = params8[6]; P.v_inter = params8[7];
    emu_launch((unsigned)grid, 1, block, [&] {
        sub_compute_likelihood(obs, s, sub_index, list_rep, list_uniq, collector, (int2*)dispatcher, (int4*)id_sub, (float3*)len_sub,
                               (int3*)accu_sub, likelihood, curr_likelihood, &P, max_id_no_repeats, lim_repeats_vs_uniq,
                               lim_intra_repeats, max_id, n_frags_uniq, n_repeats, width, n_bins, nfpb);
    });
    return 0;
}

float emu_rippe_contacts(float s, const float* q) {
    param_simu P; P.kuhn = q[0]; P.lm = q[1]; P.c1 = q[2]; P.slope = q[3]; P.d = q[4]; P.d_max = q[5]; P.fact = q[6]; P.v_inter = q[7];
    return rippe_contacts(s, P);
}
float emu_rippe_contacts_circ(float s, float s_tot, const float* q) {
    param_simu P; P.kuhn = q[0]; P.lm = q[1]; P.c1 = q[2]; P.slope = q[3]; P.d = q[4]; P.d_max = q[5]; P.fact = q[6]; P.v_inter = q[7];
    return rippe_contacts_circ(s, s_tot, P);
}
double emu_evaluate_likelihood_double(double ex, double ob) { return evaluate_likelihood_double(ex, ob); }
float emu_factorial(float n) { return factorial(n); }

}  // extern "C"
