"""Export of the assembled genome: genome.fasta + info_frags.txt.

Restates ``level.generate_new_fasta`` (/root/reference/pyramid_sparse.py:1430-1488) called by
``simulation.export_new_fasta`` (simulation_loader.py:781-783): one record per contig whose bins are
all active, bins in position order, each bin's sequence taken from its initial contig
[start_pos, end_pos) and reverse-complemented when its orientation is -1, 61 bases per line.
"""
import numpy as np

_COMPLEMENT = str.maketrans("TAGCtagc", "ATCGATCG")      # the reference maps lower case to UPPER-case complements


def read_fasta(path):
    """name -> sequence (the first word of each header is the name)."""
    seqs, name, buf = {}, None, []
    with open(path) as h:
        for line in h:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(buf)
                name, buf = line[1:].split()[0], []
            elif line:
                buf.append(line)
    if name is not None:
        seqs[name] = "".join(buf)
    return seqs


def write_fasta(path, seqs, width=60):
    with open(path, "w") as h:
        for name, s in seqs.items():
            h.write(">%s\n" % name)
            for i in range(0, len(s), width):
                h.write(s[i:i + width] + "\n")


def generate_new_fasta(vect_frags, level, contig_names, sequences, new_fasta, info_frags):
    """``vect_frags``: object or dict with id_c, pos, ori, activ, id_d arrays (sampler.gpu_vect_frags after
    copy_from_gpu()); ``level``: graal_b200.level.PyramidLevel of the sampled level; ``contig_names[c-1]``:
    name of initial contig c; ``sequences``: name -> str."""
    get = (lambda k: vect_frags[k]) if isinstance(vect_frags, dict) else (lambda k: getattr(vect_frags, k))
    id_c_frag, pos_frag, ori_frag, activ_frag, id_d = (np.asarray(get(k)) for k in ("id_c", "pos", "ori", "activ", "id_d"))
    out_seq, ok = {}, []
    with open(info_frags, "w") as hi:
        for id_cont in np.unique(id_c_frag):
            list_frags = np.nonzero(id_c_frag == id_cont)[0]
            if not np.all(activ_frag[list_frags] == 1):
                continue
            ok.append(id_cont)
            hi.write("%s\n" % (">3C-assembly|contig_" + str(id_cont)))
            hi.write("%s\t%s\t%s\t%s\t%s\n" % ("init_contig", "id_frag", "orientation", "start", "end"))
            ordered = list_frags[np.argsort(pos_frag[list_frags], kind="stable")]
            parts = []
            for f in ordered:
                ori = int(ori_frag[f])
                init_frag_id = int(id_d[f])
                init_contig = contig_names[int(level.contig_id[init_frag_id]) - 1]
                start_bp, end_bp = int(level.start_pos[init_frag_id]), int(level.end_pos[init_frag_id])
                seq = sequences[init_contig][start_bp:end_bp]
                if ori == -1:
                    seq = seq[::-1].translate(_COMPLEMENT)
                hi.write("%s\t%s\t%s\t%s\t%s\n" % (init_contig, init_frag_id, ori, start_bp, end_bp))
                parts.append(seq)
            out_seq[id_cont] = "".join(parts)
    with open(new_fasta, "w") as hf:
        for id_cont in ok:
            cont_seq = out_seq[id_cont]
            hf.write("%s\n" % (">3C-assembly|contig_" + str(id_cont)))
            len_line, len_seq = 61, len(cont_seq)
            cuts = list(range(0, len_seq, len_line))
            for i in range(1, len(cuts)):
                hf.write("%s\n" % cont_seq[cuts[i - 1]:cuts[i]])
            if cuts and cuts[-1] != len_seq - 1:        # reference quirk: a 1-base tail is dropped
                hf.write("%s\n" % cont_seq[cuts[-1]:])
    return out_seq
