// Structure mutations as pure per-bin functions.
//
// The reference implements each mutation as one kernel that copies a slot while patching the
// bins of one or two contigs (kernels3.cu:239-2070).  Here every mutation is a function
//     new_state_of_bin_i = op(old_state_of_bin_i, i, pivot states, scalars)
// so that chains (pop_out -> pop_in, split -> split -> paste) compose in registers and the 13
// candidates of a proposal are produced by ONE pass over the current slot (k_build_candidates),
// without materialising the pop / trans1 / trans2 intermediates or reducing max(id_c) in between.
#pragma once
#include <stdint.h>

struct Bin {
    int pos, id_c, start_bp, len_bp, circ, prev, next, l_cont, l_cont_bp, ori, rep, activ, id_d;
};

enum { F_POS = 0, F_ID_C, F_START_BP, F_LEN_BP, F_CIRC, F_ID, F_PREV, F_NEXT, F_L_CONT, F_L_CONT_BP,
       F_ORI, F_REP, F_ACTIV, F_ID_D, N_FIELDS };

__device__ __forceinline__ Bin load_bin(const int* __restrict__ s, int ld, int i) {
    Bin b;
    b.pos = s[F_POS * ld + i];           b.id_c = s[F_ID_C * ld + i];
    b.start_bp = s[F_START_BP * ld + i]; b.len_bp = s[F_LEN_BP * ld + i];
    b.circ = s[F_CIRC * ld + i];         b.prev = s[F_PREV * ld + i];
    b.next = s[F_NEXT * ld + i];         b.l_cont = s[F_L_CONT * ld + i];
    b.l_cont_bp = s[F_L_CONT_BP * ld + i]; b.ori = s[F_ORI * ld + i];
    b.rep = s[F_REP * ld + i];           b.activ = s[F_ACTIV * ld + i];
    b.id_d = s[F_ID_D * ld + i];
    return b;
}

__device__ __forceinline__ void store_bin(int* __restrict__ s, int ld, int i, const Bin& b) {
    s[F_POS * ld + i] = b.pos;           s[F_ID_C * ld + i] = b.id_c;
    s[F_START_BP * ld + i] = b.start_bp; s[F_LEN_BP * ld + i] = b.len_bp;
    s[F_CIRC * ld + i] = b.circ;         s[F_ID * ld + i] = i;
    s[F_PREV * ld + i] = b.prev;         s[F_NEXT * ld + i] = b.next;
    s[F_L_CONT * ld + i] = b.l_cont;     s[F_L_CONT_BP * ld + i] = b.l_cont_bp;
    s[F_ORI * ld + i] = b.ori;           s[F_REP * ld + i] = b.rep;
    s[F_ACTIV * ld + i] = b.activ;       s[F_ID_D * ld + i] = b.id_d;
}

// flip_frag (kernels3.cu:239-279)
__device__ __forceinline__ Bin op_flip(Bin s, int i, int id_f) {
    if (i == id_f) s.ori = s.ori * -1;
    return s;
}

// swap_activity_frag (kernels3.cu:283-326): only a repeat copy toggles
__device__ __forceinline__ Bin op_swap_activity(Bin s, int i, int id_f, int max_id) {
    if (i == id_f && s.rep == 1) {
        if (s.activ == 1) { s.activ = 0; }
        else if (s.activ == 0) { s.activ = 1; s.id_c = max_id + 1; }
        else { s.activ = 0; s.id_c = 0; }          // reference arithmetic for activ outside {0,1}
    }
    return s;
}

// pop_out_frag (kernels3.cu:329-563).  P = source state of the ejected bin.
__device__ __forceinline__ Bin op_pop_out(Bin s, int i, const Bin& P, int max_id) {
    const int lc = P.l_cont;
    if (lc < 2 || s.id_c != P.id_c) return s;
    if (s.pos == P.pos) {
        s.pos = 0; s.id_c = max_id + 1; s.start_bp = 0; s.circ = 0; s.ori = 1;
        s.prev = -1; s.next = -1; s.l_cont = 1; s.l_cont_bp = s.len_bp;
        return s;
    }
    s.l_cont -= 1;
    s.l_cont_bp -= P.len_bp;
    const bool after = s.pos > P.pos;
    if (lc == 2) {
        s.circ = 0; s.prev = -1; s.next = -1;
    } else if (!after) {
        if (i == P.next && P.circ == 1) s.prev = P.prev;
        if (s.pos == P.pos - 1) s.next = P.next;
    } else {
        if (s.pos == P.pos + 1) s.prev = P.prev;
        if (i == P.prev && P.circ == 1) s.next = P.next;
    }
    if (after) { s.pos -= 1; s.start_bp -= P.len_bp; }
    return s;
}

// does pop_out create contig id max_id+1?  (the ga.max after pop_out, cuda_lib_gl.py:857)
__device__ __forceinline__ int pop_out_new_ids(const Bin& P) { return P.l_cont >= 2 ? 1 : 0; }

// pop_in_frag_1..4 (kernels3.cu:565-1448).  Pp / Pi = source states of the popped bin and of the
// insertion partner; kind 1 = split insert @ left, 2 = split insert @ right, 3 = insert right of
// id_f_ins, 4 = insert left of id_f_ins.
__device__ __forceinline__ Bin op_pop_in(int kind, Bin s, int i, const Bin& Pp, const Bin& Pi,
                                         int id_f_pop, int id_f_ins, int max_id, int ori) {
    if (!(Pi.activ == 1 && Pp.activ == 1)) return s;
    const int new_id = max_id + 1;
    const int len_p = Pp.len_bp;
    const int end_i = Pi.start_bp + Pi.len_bp;
    const bool lin = Pi.circ == 0;
    if (i == id_f_pop) {
        s.len_bp = len_p; s.ori = ori;
        if (kind == 1) {
            s.pos = 0; s.start_bp = 0; s.circ = 0; s.prev = -1; s.next = id_f_ins;
            if (lin) { s.id_c = new_id; s.l_cont = Pi.l_cont - Pi.pos + 1; s.l_cont_bp = Pi.l_cont_bp - Pi.start_bp + len_p; }
            else     { s.id_c = Pi.id_c; s.l_cont = Pi.l_cont + 1; s.l_cont_bp = Pi.l_cont_bp + len_p; }
        } else if (kind == 2) {
            s.id_c = Pi.id_c; s.circ = 0; s.prev = id_f_ins; s.next = -1;
            if (lin) { s.pos = Pi.pos + 1; s.start_bp = end_i; s.l_cont = Pi.pos + 2; s.l_cont_bp = end_i + len_p; }
            else {
                s.pos = (Pi.l_cont - (Pi.pos + 1)) + Pi.pos + 1;
                s.start_bp = (Pi.l_cont_bp - end_i) + end_i;
                s.l_cont = Pi.l_cont + 1; s.l_cont_bp = Pi.l_cont_bp + len_p;
            }
        } else if (kind == 3) {
            s.pos = Pi.pos + 1; s.id_c = Pi.id_c; s.start_bp = end_i; s.circ = Pi.circ;
            s.prev = id_f_ins; s.next = Pi.next; s.l_cont = Pi.l_cont + 1; s.l_cont_bp = Pi.l_cont_bp + len_p;
        } else {
            s.pos = Pi.pos; s.id_c = Pi.id_c; s.start_bp = Pi.start_bp; s.circ = Pi.circ;
            s.prev = Pi.prev; s.next = id_f_ins; s.l_cont = Pi.l_cont + 1; s.l_cont_bp = Pi.l_cont_bp + len_p;
        }
        return s;
    }
    if (s.id_c != Pi.id_c) return s;
    const int rel = (s.pos < Pi.pos) ? -1 : (s.pos == Pi.pos ? 0 : 1);
    if (kind == 3 || kind == 4) {
        s.circ = Pi.circ; s.l_cont = Pi.l_cont + 1; s.l_cont_bp = Pi.l_cont_bp + len_p;
        if (kind == 3) {
            if (rel < 0) { if (i == Pi.next && Pi.circ == 1) s.prev = id_f_pop; }
            else if (rel == 0) { s.ori = Pi.ori; s.next = id_f_pop; }
            else { if (s.pos == Pi.pos + 1) s.prev = id_f_pop; s.pos += 1; s.start_bp += len_p; }
        } else {
            if (rel < 0) { if (s.pos == Pi.pos - 1) s.next = id_f_pop; }
            else if (rel == 0) { s.pos = Pi.pos + 1; s.start_bp = Pi.start_bp + len_p; s.ori = Pi.ori; s.prev = id_f_pop; s.next = Pi.next; }
            else { s.pos += 1; s.start_bp += len_p; }
        }
        return s;
    }
    s.circ = 0;
    if (kind == 1) {
        if (lin) {
            const int lc_new = Pi.l_cont - Pi.pos + 1, lcb_new = Pi.l_cont_bp - Pi.start_bp + len_p;
            if (rel < 0) { if (s.pos == Pi.pos - 1) s.next = -1; s.l_cont = Pi.pos; s.l_cont_bp = Pi.start_bp; }
            else if (rel == 0) {
                s.pos = 1; s.id_c = new_id; s.start_bp = len_p; s.ori = Pi.ori; s.prev = id_f_pop; s.next = Pi.next;
                s.l_cont = lc_new; s.l_cont_bp = lcb_new;
            } else {
                s.pos = s.pos - Pi.pos + 1; s.id_c = new_id; s.start_bp = s.start_bp - Pi.start_bp + len_p;
                s.l_cont = lc_new; s.l_cont_bp = lcb_new;
            }
        } else {
            if (rel < 0) {
                if (s.pos == Pi.pos - 1) s.next = -1;
                s.pos = Pi.l_cont - Pi.pos + s.pos + 1;
                s.start_bp = Pi.l_cont_bp - Pi.start_bp + s.start_bp + len_p;
            } else if (rel == 0) {
                s.pos = 1; s.start_bp = len_p; s.len_bp = Pi.len_bp; s.ori = Pi.ori; s.prev = id_f_pop; s.next = Pi.next;
            } else {
                if (i == Pi.prev) s.next = -1;
                s.pos = s.pos - Pi.pos + 1; s.start_bp = s.start_bp - Pi.start_bp + len_p;
            }
            s.l_cont = Pi.l_cont + 1; s.l_cont_bp = Pi.l_cont_bp + len_p;
        }
    } else {  // kind == 2
        if (lin) {
            if (rel <= 0) {
                if (rel == 0) { s.ori = Pi.ori; s.prev = Pi.prev; s.next = id_f_pop; }
                s.l_cont = Pi.pos + 2; s.l_cont_bp = end_i + len_p;
            } else {
                if (s.pos == Pi.pos + 1) s.prev = -1;
                s.pos = s.pos - (Pi.pos + 1); s.id_c = new_id; s.start_bp = s.start_bp - end_i;
                s.l_cont = Pi.l_cont - (Pi.pos + 1); s.l_cont_bp = Pi.l_cont_bp - end_i;
            }
        } else {
            const int sh_pos = Pi.l_cont - (Pi.pos + 1), sh_bp = Pi.l_cont_bp - end_i;
            if (rel < 0) { if (i == Pi.next) s.prev = -1; s.pos = sh_pos + s.pos; s.start_bp = sh_bp + s.start_bp; }
            else if (rel == 0) { s.pos = sh_pos + Pi.pos; s.start_bp = sh_bp + Pi.start_bp; s.len_bp = Pi.len_bp; s.prev = Pi.prev; s.next = id_f_pop; }
            else { if (s.pos == Pi.pos + 1) s.prev = -1; s.pos = s.pos - (Pi.pos + 1); s.start_bp = s.start_bp - end_i; }
            s.l_cont = Pi.l_cont + 1; s.l_cont_bp = Pi.l_cont_bp + len_p;
        }
    }
    return s;
}

// split_contig (kernels3.cu:1451-1784).  Pc = source state of the cut bin.
__device__ __forceinline__ Bin op_split(Bin s, int i, const Bin& Pc, int upstream, int max_id) {
    if (!(Pc.activ == 1 && Pc.l_cont > 1) || s.id_c != Pc.id_c) return s;
    const int new_id = max_id + 1;
    const int end_c = Pc.start_bp + Pc.len_bp;
    const int rel = (s.pos < Pc.pos) ? -1 : (s.pos == Pc.pos ? 0 : 1);
    s.circ = 0;
    if (Pc.circ == 0) {
        if (upstream == 1) {
            if (rel < 0) { if (s.pos == Pc.pos - 1) s.next = -1; s.l_cont = Pc.pos; s.l_cont_bp = Pc.start_bp; }
            else {
                if (rel == 0) { s.len_bp = Pc.len_bp; s.prev = -1; s.next = Pc.next; }
                s.pos = s.pos - Pc.pos; s.start_bp = (rel == 0) ? 0 : s.start_bp - Pc.start_bp; s.id_c = new_id;
                s.l_cont = Pc.l_cont - Pc.pos; s.l_cont_bp = Pc.l_cont_bp - Pc.start_bp;
            }
        } else {
            if (rel <= 0) {
                if (rel == 0) { s.start_bp = Pc.start_bp; s.len_bp = Pc.len_bp; s.prev = Pc.prev; s.next = -1; }
                s.l_cont = Pc.pos + 1; s.l_cont_bp = end_c;
            } else {
                if (s.pos == Pc.pos + 1) s.prev = -1;
                s.pos = s.pos - (Pc.pos + 1); s.id_c = new_id; s.start_bp = s.start_bp - end_c;
                s.l_cont = Pc.l_cont - (Pc.pos + 1); s.l_cont_bp = Pc.l_cont_bp - end_c;
            }
        }
    } else {   // circular contig: linearised at the cut, keeps its id and its lengths
        if (upstream == 1) {
            if (rel < 0) {
                if (s.pos == Pc.pos - 1) s.next = -1;
                s.pos = Pc.l_cont - Pc.pos + s.pos; s.start_bp = Pc.l_cont_bp - Pc.start_bp + s.start_bp;
            } else if (rel == 0) { s.pos = 0; s.start_bp = 0; s.len_bp = Pc.len_bp; s.prev = -1; s.next = Pc.next; }
            else { if (i == Pc.prev) s.next = -1; s.pos = s.pos - Pc.pos; s.start_bp = s.start_bp - Pc.start_bp; }
        } else {
            const int sh_pos = Pc.l_cont - (Pc.pos + 1), sh_bp = Pc.l_cont_bp - end_c;
            if (rel < 0) { if (i == Pc.next) s.prev = -1; s.pos = sh_pos + s.pos; s.start_bp = sh_bp + s.start_bp; }
            else if (rel == 0) { s.pos = sh_pos + s.pos; s.start_bp = sh_bp + Pc.start_bp; s.len_bp = Pc.len_bp; s.prev = Pc.prev; s.next = -1; }
            else { if (s.pos == Pc.pos + 1) s.prev = -1; s.pos = s.pos - (Pc.pos + 1); s.start_bp = s.start_bp - end_c; }
        }
        s.l_cont = Pc.l_cont; s.l_cont_bp = Pc.l_cont_bp;
    }
    return s;
}

// does split create contig id max_id+1?  (the ga.max after split, cuda_lib_gl.py:934,943)
__device__ __forceinline__ int split_new_ids(const Bin& Pc, int upstream) {
    if (!(Pc.activ == 1 && Pc.l_cont > 1) || Pc.circ != 0) return 0;
    return (upstream == 1 || Pc.pos < Pc.l_cont - 1) ? 1 : 0;
}

// paste_contigs (kernels3.cu:1786-2070).  Returns false when the reference kernel writes nothing
// for this bin (same contig, not end-to-end: persistent destination).
__device__ __forceinline__ bool op_paste(Bin& s, int i, const Bin& PA, const Bin& PB, int id_fA, int id_fB) {
    if (!(PA.activ == 1 && PB.activ == 1)) return true;
    if (PA.id_c != PB.id_c) {
        const int lc = PA.l_cont + PB.l_cont, lcb = PA.l_cont_bp + PB.l_cont_bp;
        if (s.id_c == PA.id_c) {
            if (PA.pos == 0) {      // reverse contig A
                const int p = s.pos, nx = s.next, pv = s.prev;
                s.start_bp = PA.l_cont_bp - (s.start_bp + s.len_bp);
                s.pos = PA.l_cont - (p + 1);
                s.ori = s.ori * -1;
                s.prev = (p == PA.l_cont - 1) ? -1 : nx;
                s.next = (p == PA.pos) ? id_fB : pv;
            } else {
                if (s.pos == PA.pos) s.next = id_fB;
            }
            s.id_c = PA.id_c; s.circ = 0; s.l_cont = lc; s.l_cont_bp = lcb;
        } else if (s.id_c == PB.id_c) {
            if (PB.pos == 0) {
                if (s.pos == PB.pos) s.prev = id_fA;
                s.pos = PA.l_cont + s.pos; s.start_bp = PA.l_cont_bp + s.start_bp;
            } else {               // reverse contig B
                const int p = s.pos, nx = s.next, pv = s.prev;
                s.start_bp = PA.l_cont_bp + (PB.l_cont_bp - (s.start_bp + s.len_bp));
                s.pos = PA.l_cont + (PB.l_cont - (p + 1));
                s.ori = s.ori * -1;
                s.prev = (p == PB.pos) ? id_fA : nx;
                s.next = (p == 0) ? -1 : pv;
            }
            s.id_c = PA.id_c; s.circ = 0; s.l_cont = lc; s.l_cont_bp = lcb;
        }
        return true;
    }
    if (s.id_c != PA.id_c) return true;
    if (PA.pos == 0 && PB.pos == PA.l_cont - 1) {
        if (s.pos == PA.pos) s.prev = id_fB;
        if (s.pos == PA.l_cont - 1) s.next = id_fA;
    } else if (PA.pos == PA.l_cont - 1 && PB.pos == 0) {
        if (s.pos == PB.pos) s.prev = id_fA;
        if (s.pos == PA.l_cont - 1) s.next = id_fB;
    } else {
        return false;
    }
    s.circ = 1; s.l_cont = PA.l_cont; s.l_cont_bp = PA.l_cont_bp;
    return true;
}
