// C entry points over the host logic, for the CPU test-suite (ctypes): the partition builder on flat
// arrays and a dump of what the parsers keep. Not part of the product ABI (include/hsgpu.h).
#include <fstream>

#include "hs_io.h"
#include "hs_partition.h"
#include "hs_sepreads.h"

using namespace hs;

struct PartSet {
    std::vector<Partition> parts;
};

extern "C" {

void* hshost_build_partitions(int n_cols, const int64_t* off, const uint32_t* read_idx, const uint8_t* code,
                              const int32_t* pos, const uint8_t* ref_base, const uint8_t* second_base, float mean_error) {
    std::vector<Column> cols(n_cols);
    for (int i = 0; i < n_cols; i++) {
        cols[i].pos = pos[i];
        cols[i].ref_base = ref_base[i];
        cols[i].second_base = second_base[i];
        cols[i].readIdxs.assign(read_idx + off[i], read_idx + off[i + 1]);
        cols[i].content.assign(code + off[i], code + off[i + 1]);
    }
    PartSet* ps = new PartSet();
    build_partitions(cols, mean_error, ps->parts);
    return ps;
}
int hshost_parts_count(void* h) { return (int)((PartSet*)h)->parts.size(); }
int hshost_part_size(void* h, int p) { return (int)((PartSet*)h)->parts[p].readIdx.size(); }
void hshost_part_get(void* h, int p, int32_t* read_idx, int16_t* state, int32_t* more, int32_t* less, int32_t* left_right) {
    const Partition& pt = ((PartSet*)h)->parts[p];
    for (size_t i = 0; i < pt.readIdx.size(); i++) {
        read_idx[i] = pt.readIdx[i];
        state[i] = pt.state[i];
        more[i] = pt.more[i];
        less[i] = pt.less[i];
    }
    left_right[0] = pt.pos_left;
    left_right[1] = pt.pos_right;
}
void hshost_parts_free(void* h) { delete (PartSet*)h; }

// create_read_graph_low_memory of the host path on flat SNP columns: CSR over the masked reads, local indices;
// also reports whether the device path may replace it (every read on consecutive SNP columns)
int64_t hshost_read_graph_low_memory(int n_reads, int n_snps, const int64_t* snp_off, const uint32_t* idx, const uint8_t* code,
                                     const uint8_t* rb, const uint8_t* sb, int m, const int32_t* masked, float error_rate,
                                     int64_t* adj_off, int32_t* adj, int* consecutive) {
    ColContig c;
    c.read_lines.resize((size_t)n_reads);
    c.snps.resize((size_t)n_snps);
    for (int s = 0; s < n_snps; s++) {
        Column& col = c.snps[s];
        col.pos = s;
        col.ref_base = rb[s];
        col.second_base = sb[s];
        col.readIdxs.assign(idx + snp_off[s], idx + snp_off[s + 1]);
        col.content.assign(code + snp_off[s], code + snp_off[s + 1]);
    }
    if (consecutive) *consecutive = low_memory_counts_are_contig_counts(c) ? 1 : 0;
    std::vector<char> mask((size_t)n_reads, 0);
    std::vector<int> local((size_t)n_reads, -1);
    for (int i = 0; i < m; i++) {
        mask[masked[i]] = 1;
        local[masked[i]] = i;
    }
    ReadGraph g;
    create_read_graph_low_memory(c.snps, mask, g, error_rate);
    int64_t n = 0;
    for (int i = 0; i < m; i++) {
        adj_off[i] = n;
        for (int e = g.off[masked[i]]; e < g.off[masked[i] + 1]; e++) {
            if (adj) adj[n] = local[g.nbr[e]];
            n++;
        }
    }
    adj_off[m] = n;
    return n;
}

float hshost_chi_square(int n00, int n01, int n10, int n11) {
    Distance d;
    d.n00 = n00; d.n01 = n01; d.n10 = n10; d.n11 = n11;
    return chi_square(d);
}

// CONTIG / READ lines exactly as output_files prints them (depth left out), one block per contig in GFA order
int hshost_parse_dump(const char* gfa, const char* reads, const char* sam, int amplicon, const char* out_path) {
    try {
        Store st;
        parse_reads(reads, st);
        parse_assembly(gfa, st);
        parse_sam(sam, st, amplicon != 0);
        std::ofstream out(out_path);
        std::ifstream rf(reads);
        std::vector<std::string> seqs;
        for (int64_t c : st.contigs) {
            out << "CONTIG\t" << st.seqs[c].name << "\t" << st.seqs[c].sequence.size() << "\n";
            load_read_sequences(rf, st, c, seqs);
            size_t n = 0;
            for (int64_t id : st.seqs[c].alns) {
                const Alignment& a = st.alns[id];
                out << "READ\t" << st.seqs[a.read].name << "\t" << a.pos_1_1 << "\t" << a.pos_1_2 << "\t" << a.pos_2_1 << "\t"
                    << a.pos_2_2 << "\t" << a.strand << "\t" << seqs[n++].size() << "\n";
            }
        }
        return 0;
    } catch (...) {
        return 1;
    }
}

}  // extern "C"

// ---- .col sidecar (hs_colbin.h) ----------------------------------------------------------------------------------
#include <chrono>
#include <sstream>

#include "hs_colbin.h"

namespace {
uint64_t mix(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) {
        h ^= b[i];
        h *= 1099511628211ull;
    }
    return h;
}
}  // namespace

extern "C" {

// Everything parse_column_file leaves behind, folded into one number, plus counts: out = {digest, contigs, snps,
// cells, route (1 = the sidecar was used)}. route_wanted: 0 = text only, 1 = sidecar only (fails with -1 if there is no
// usable one), 2 = what HS_separate_reads does. Returns the seconds the parse took, < 0 on failure.
double hshost_col_digest(const char* col_path, int max_coverage, float rarest, int route_wanted, uint64_t* out) {
    std::vector<ColContig> contigs;
    const auto t0 = std::chrono::steady_clock::now();
    int route = 0;
    if (route_wanted == 0) parse_column_text(col_path, contigs, max_coverage, rarest);
    else if (read_col_sidecar(col_path, contigs, max_coverage, rarest)) route = 1;
    else if (route_wanted == 1) return -1.0;
    else parse_column_text(col_path, contigs, max_coverage, rarest);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t h = 1469598103934665603ull, n_snps = 0, n_cells = 0;
    for (const ColContig& c : contigs) {
        h = mix(h, c.line.data(), c.line.size());
        h = mix(h, &c.length, sizeof(c.length));
        h = mix(h, &c.coverage, sizeof(c.coverage));
        const uint64_t nr = c.read_lines.size(), nl = c.limits.size(), ns = c.snps.size();
        h = mix(h, &nr, 8);
        h = mix(h, &nl, 8);
        h = mix(h, &ns, 8);
        for (const std::string& r : c.read_lines) h = mix(h, r.data(), r.size() + 1);
        for (const auto& l : c.limits) {
            h = mix(h, &l.first, sizeof(int));
            h = mix(h, &l.second, sizeof(int));
        }
        for (const Column& s : c.snps) {
            const uint64_t m = s.readIdxs.size(), k = s.content.size();
            h = mix(h, &s.pos, sizeof(int));
            h = mix(h, &s.ref_base, 1);
            h = mix(h, &s.second_base, 1);
            h = mix(h, &m, 8);
            h = mix(h, &k, 8);
            h = mix(h, s.readIdxs.data(), 4 * m);
            h = mix(h, s.content.data(), k);
            n_cells += k;
        }
        n_snps += ns;
    }
    out[0] = h;
    out[1] = contigs.size();
    out[2] = n_snps;
    out[3] = n_cells;
    out[4] = (uint64_t)route;
    return dt;
}

// a .col text (any writer's) through write_outputs again: col_out must reproduce it block for block, and the sidecar
// write_outputs leaves next to col_out must parse to the same structures as the text
int hshost_rewrite_col(const char* col_in, const char* col_out, const char* vcf_out) {
    try {
        std::vector<ColContig> contigs;
        parse_column_text(col_in, contigs, 2147483647, 0.0f);
        Store st;
        std::unordered_map<int, std::vector<Column>> variants;
        for (ColContig& c : contigs) {
            std::istringstream cl(c.line);
            std::string tag, name, len, cov;
            cl >> tag >> name >> len >> cov;
            SeqRec contig;
            contig.name = name;
            contig.sequence.assign((size_t)c.length, 'A');
            contig.depth = std::strtof(cov.c_str(), nullptr);
            const int64_t ci = (int64_t)st.seqs.size();
            st.seqs.push_back(std::move(contig));
            st.contigs.push_back(ci);
            for (const std::string& rl : c.read_lines) {
                std::istringstream is(rl);
                std::string rname;
                Alignment a;
                int strand = 0;
                is >> tag >> rname >> a.pos_1_1 >> a.pos_1_2 >> a.pos_2_1 >> a.pos_2_2 >> strand;
                a.strand = strand != 0;
                a.contig = ci;
                a.read = (int64_t)st.seqs.size();
                SeqRec read;
                read.name = rname;
                st.seqs.push_back(std::move(read));
                st.seqs[ci].alns.push_back((int64_t)st.alns.size());
                st.alns.push_back(std::move(a));
            }
            variants[(int)ci] = std::move(c.snps);
        }
        const auto t0 = std::chrono::steady_clock::now();
        write_outputs(st, variants, col_out, vcf_out);
        if (std::getenv("HS_TIMING"))
            fprintf(stderr, "[hs timing] write_outputs %.3f s\n",
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return 0;
    } catch (...) {
        return 1;
    }
}

int hshost_col_sidecar_enabled() { return col_sidecar_enabled() ? 1 : 0; }

// cigar_ops of the host parsers: number of ops, the ops in out (capacity entries at most are written)
int64_t hshost_cigar_ops(const char* cigar, uint32_t* out, int64_t capacity) {
    std::vector<uint32_t> ops;
    cigar_ops(cigar, ops);
    for (size_t k = 0; k < ops.size() && (int64_t)k < capacity; k++) out[k] = ops[k];
    return (int64_t)ops.size();
}

void hshost_pack_bases_2bit(const char* seq, int64_t n, uint32_t* out) { pack_bases_2bit(seq, n, out); }

}  // extern "C"
