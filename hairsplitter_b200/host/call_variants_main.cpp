// HS_call_variants -- drop-in for the reference executable of the same name (src/call_variants.cpp:1215-1386,
// called by hairsplitter.py with 11 positional arguments). Parsing, partition building and the writers run
// on the host; the pileup (generate_msa), the per-column allele ranking (call_variants) and the
// partition x column filtering (loops 3+4 of keep_only_robust_variants) run on the GPU through the
// C ABI of libhsgpu (include/hsgpu.h). Contigs are sharded over the visible GPUs, heaviest first.
#include <omp.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <functional>
#include <future>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hsgpu.h"
#include "hs_io.h"
#include "hs_partition.h"

using namespace hs;

// HS_TIMING=1 prints the wall time of every phase on stderr
static double g_t0 = 0;
static bool g_timing = false;
static void phase(const char* name) {
    if (!g_timing) return;
    const double t = omp_get_wtime();
    if (g_t0 > 0) fprintf(stderr, "[hs timing] %-28s %8.3f s\n", name, t - g_t0);
    g_t0 = t;
}

#define GPU_CHECK(ctx, call)                                                                        \
    do {                                                                                            \
        const int _rc = (call);                                                                     \
        if (_rc != HSGPU_OK) {                                                                      \
            std::cout << "ERROR: " #call " failed (" << _rc << "): " << hsgpu_last_error(ctx) << std::endl; \
            std::exit(1);                                                                           \
        }                                                                                           \
    } while (0)

// the 16-bit CIGAR form of hsgpu_pileup_input (ops longer than 4095 become several ops of the same kind)
static void compact_ops(const std::vector<uint32_t>& ops, std::vector<uint16_t>& out) {
    out.clear();
    for (uint32_t op : ops) {
        uint32_t n = op >> 4;
        const uint16_t ty = (uint16_t)(op & 15);
        while (n > 4095) {
            out.push_back((uint16_t)((4095u << 4) | ty));
            n -= 4095;
        }
        out.push_back((uint16_t)((n << 4) | ty));
    }
}

struct ContigResult {
    float mean_distance = 0;
    float depth = 0;
    std::vector<Column> merged;
};

static void fetch_columns(hsgpu_ctx* ctx, hsgpu_pileup* pu, int c, const std::vector<int32_t>& pos,
                          const std::vector<uint8_t>& k0, const std::vector<uint8_t>& k1, std::vector<Column>& out) {
    out.clear();
    const int n = (int)pos.size();
    if (n == 0) return;
    std::vector<int64_t> off(n + 1, 0);
    // first call sizes the cell buffers
    int rc = hsgpu_pileup_extract_columns(pu, c, n, pos.data(), 0, off.data(), nullptr, nullptr);
    if (rc != HSGPU_OK && rc != HSGPU_ERR_CAPACITY) GPU_CHECK(ctx, rc);
    std::vector<uint32_t> idx((size_t)std::max<int64_t>(off[n], 1));
    std::vector<uint8_t> code((size_t)std::max<int64_t>(off[n], 1));
    GPU_CHECK(ctx, hsgpu_pileup_extract_columns(pu, c, n, pos.data(), off[n], off.data(), idx.data(), code.data()));
    out.resize(n);
    for (int i = 0; i < n; i++) {
        Column& col = out[i];
        col.pos = pos[i];
        col.ref_base = k0[pos[i]];
        col.second_base = k1[pos[i]];
        col.readIdxs.assign(idx.begin() + off[i], idx.begin() + off[i + 1]);
        col.content.assign(code.begin() + off[i], code.begin() + off[i + 1]);
    }
}

// one batch of contigs on one GPU
// (get_ctx waits for the context that is being created in the background: the reads are loaded and packed first)
static void process_batch(const std::function<hsgpu_ctx*()>& get_ctx, const Store& st, const MappedFile& reads_map,
                          const std::vector<int>& batch, float auto_threshold, std::vector<ContigResult>& results) {
    const int nc = (int)batch.size();
    std::vector<int32_t> contig_len(nc), read_len, read_start;
    std::vector<int64_t> contig_word_off(nc + 1, 0), contig_read_off(nc + 1, 0), read_word_off(1, 0), cigar_off(1, 0);
    std::vector<uint32_t> contig_bases, read_bases;
    std::vector<uint16_t> cigar;
    std::vector<uint8_t> read_strand;
    {
        // the reads of every contig: views of their sequence lines in the mapped reads file (no copies) and CIGAR ops,
        // contigs in parallel, then one pass for the offsets, then the 2-bit packing straight from the mapping
        std::vector<std::vector<std::pair<const char*, size_t>>> seqs(nc);
        std::vector<std::vector<std::vector<uint16_t>>> ops(nc);
#pragma omp parallel
        {
#pragma omp for schedule(dynamic, 1)
            for (int b = 0; b < nc; b++) {
                const SeqRec& contig = st.seqs[st.contigs[batch[b]]];
                view_read_sequences(reads_map, st, st.contigs[batch[b]], seqs[b]);
                ops[b].resize(contig.alns.size());
                std::vector<uint32_t> wide;
                for (size_t n = 0; n < contig.alns.size(); n++) {
                    cigar_ops(st.alns[contig.alns[n]].cigar, wide);
                    compact_ops(wide, ops[b][n]);
                }
            }
        }
        std::vector<int64_t> first_read(nc + 1, 0);
        for (int b = 0; b < nc; b++) {
            const SeqRec& contig = st.seqs[st.contigs[batch[b]]];
            contig_len[b] = (int32_t)contig.sequence.size();
            contig_word_off[b + 1] = contig_word_off[b] + ((int64_t)contig.sequence.size() + 15) / 16;
            for (size_t n = 0; n < contig.alns.size(); n++) {
                const Alignment& a = st.alns[contig.alns[n]];
                read_word_off.push_back(read_word_off.back() + ((int64_t)seqs[b][n].second + 15) / 16);
                read_len.push_back((int32_t)seqs[b][n].second);
                cigar_off.push_back(cigar_off.back() + (int64_t)ops[b][n].size());
                read_start.push_back(a.pos_2_1);
                read_strand.push_back(a.strand ? 1 : 0);
            }
            contig_read_off[b + 1] = (int64_t)read_len.size();
        }
        contig_bases.resize((size_t)contig_word_off[nc]);
        read_bases.resize((size_t)read_word_off.back());
        cigar.resize((size_t)cigar_off.back());
#pragma omp parallel for schedule(dynamic, 1)
        for (int b = 0; b < nc; b++) {
            const SeqRec& contig = st.seqs[st.contigs[batch[b]]];
            pack_bases_2bit(contig.sequence.data(), (int64_t)contig.sequence.size(), contig_bases.data() + contig_word_off[b]);
            for (size_t n = 0; n < contig.alns.size(); n++) {
                const int64_t r = contig_read_off[b] + (int64_t)n;
                pack_bases_2bit(seqs[b][n].first, (int64_t)seqs[b][n].second, read_bases.data() + read_word_off[r]);
                std::copy(ops[b][n].begin(), ops[b][n].end(), cigar.begin() + cigar_off[r]);
            }
        }
    }
    if (contig_bases.empty()) contig_bases.push_back(0);
    if (read_bases.empty()) read_bases.push_back(0);
    if (cigar.empty()) cigar.push_back(0);
    hsgpu_pileup_input in;
    std::memset(&in, 0, sizeof(in));
    in.n_contigs = nc;
    in.contig_len = contig_len.data();
    in.contig_bases = contig_bases.data();
    in.contig_word_off = contig_word_off.data();
    in.contig_read_off = contig_read_off.data();
    in.n_reads = (int64_t)read_len.size();
    in.read_bases = read_bases.data();
    in.read_word_off = read_word_off.data();
    in.read_len = read_len.data();
    in.cigar16 = cigar.data();
    in.cigar_off = cigar_off.data();
    in.read_start = read_start.data();
    in.read_strand = read_strand.data();
    phase("  load+pack reads");
    hsgpu_ctx* ctx = get_ctx();
    phase("  wait for the CUDA context");
    hsgpu_pileup* pu = nullptr;
    GPU_CHECK(ctx, hsgpu_pileup_create(ctx, &in, &pu));
    phase("    upload");
    GPU_CHECK(ctx, hsgpu_pileup_build(pu));                            // generate_msa
    GPU_CHECK(ctx, hsgpu_column_rank(pu, nullptr, auto_threshold));    // call_variants
    std::vector<int64_t> cells(nc), dist(nc), alen(nc), depth_sum(nc);
    std::vector<int32_t> n_suspects(nc);
    GPU_CHECK(ctx, hsgpu_pileup_stats(pu, cells.data(), dist.data(), alen.data()));
    GPU_CHECK(ctx, hsgpu_column_counts(pu, n_suspects.data(), depth_sum.data()));

    phase("  gpu pileup+rank");
    // phase A: suspect columns of every contig (device -> host)
    std::vector<std::vector<Column>> suspects(nc);
    std::vector<std::vector<uint8_t>> is_auto(nc), k0(nc), k1(nc);
    std::vector<std::vector<int32_t>> suspect_pos(nc);
    for (int b = 0; b < nc; b++) {
        const int L = contig_len[b];
        k0[b].resize(std::max(L, 1));
        k1[b].resize(std::max(L, 1));
        GPU_CHECK(ctx, hsgpu_column_summary(pu, b, k0[b].data(), k1[b].data(), nullptr, nullptr));
        suspect_pos[b].resize(std::max(n_suspects[b], 1));
        is_auto[b].resize(std::max(n_suspects[b], 1));
        GPU_CHECK(ctx, hsgpu_suspects(pu, b, n_suspects[b], suspect_pos[b].data(), is_auto[b].data()));
        suspect_pos[b].resize(n_suspects[b]);
        is_auto[b].resize(n_suspects[b]);
        fetch_columns(ctx, pu, b, suspect_pos[b], k0[b], k1[b], suspects[b]);
    }
    phase("  fetch suspect columns");
    // phase B: partitions (loops 1+2 of keep_only_robust_variants), contigs in parallel on the host
    std::vector<std::vector<Partition>> parts(nc);
    std::vector<float> mean_distance(nc);
    for (int b = 0; b < nc; b++) mean_distance[b] = hsgpu_mean_distance(dist[b], alen[b]);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nc; b++) build_partitions(suspects[b], mean_distance[b], parts[b]);
    phase("  build partitions");
    // phase C: loops 3+4 on the device -- the final partitions of every contig of the batch go up in one piece and
    // one launch filters all contigs -- then the merge with the automatic SNPs (main(), :1334-1352)
    std::vector<std::vector<int64_t>> part_off(nc);
    std::vector<std::vector<int32_t>> p_idx(nc), p_more(nc), p_less(nc);
    std::vector<std::vector<int16_t>> p_state(nc);
    std::vector<hsgpu_partitions> hp(nc);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nc; b++) {
        part_off[b].assign(1, 0);
        for (const Partition& p : parts[b]) {
            p_idx[b].insert(p_idx[b].end(), p.readIdx.begin(), p.readIdx.end());
            p_state[b].insert(p_state[b].end(), p.state.begin(), p.state.end());
            p_more[b].insert(p_more[b].end(), p.more.begin(), p.more.end());
            p_less[b].insert(p_less[b].end(), p.less.begin(), p.less.end());
            part_off[b].push_back((int64_t)p_idx[b].size());
        }
        hp[b].n_parts = (int32_t)parts[b].size();
        hp[b].part_off = part_off[b].data();
        hp[b].read_idx = p_idx[b].data();
        hp[b].state = p_state[b].data();
        hp[b].more = p_more[b].data();
        hp[b].less = p_less[b].data();
    }
    GPU_CHECK(ctx, hsgpu_partitions_set(pu, hp.data()));
    std::vector<int64_t> kept_off(nc + 1, 0);
    int64_t total_cols = 0;
    for (int b = 0; b < nc; b++) total_cols += contig_len[b];
    std::vector<int32_t> kept_all((size_t)std::max<int64_t>(total_cols, 1));
    GPU_CHECK(ctx, hsgpu_robust_filter_all(pu, total_cols, kept_all.data(), kept_off.data()));
    for (int b = 0; b < nc; b++) {
        ContigResult& res = results[batch[b]];
        res.mean_distance = mean_distance[b];
        res.depth = (float)((double)depth_sum[b] / (size_t)contig_len[b]);
        std::vector<Column> filtered;
        if (!parts[b].empty()) {
            std::vector<int32_t> kept(kept_all.begin() + kept_off[b], kept_all.begin() + kept_off[b + 1]);
            fetch_columns(ctx, pu, b, kept, k0[b], k1[b], filtered);
        }
        size_t ia = 0, jf = 0;
        std::vector<const Column*> automatic;
        for (size_t i = 0; i < suspects[b].size(); i++)
            if (is_auto[b][i]) automatic.push_back(&suspects[b][i]);
        // the reference's merge stops as soon as either list is exhausted (:1337-1352)
        while (ia < automatic.size() && jf < filtered.size()) {
            if (automatic[ia]->pos < filtered[jf].pos) res.merged.push_back(*automatic[ia++]);
            else if (automatic[ia]->pos > filtered[jf].pos) res.merged.push_back(std::move(filtered[jf++]));
            else { res.merged.push_back(*automatic[ia++]); jf++; }
        }
    }
    phase("  robust filter + fetch");
    hsgpu_pileup_destroy(pu);
}

int main(int argc, char* argv[]) {
    if (argc < 12) {
        std::cout << "Usage: ./call_variants <gfa_file> <reads_file> <sam_file> <num_threads> <tmpDir> <error_rate_out> "
                     "<amplicon> <DEBUG> <file_out> <vcfFile> <automatic_snp_threshold>\n";
        return 0;
    }
    const double t_main = omp_get_wtime();
    const std::string gfa_file = argv[1], reads_file = argv[2], sam_file = argv[3];
    const int num_threads = std::stoi(argv[4]);
    const std::string error_rate_out = argv[6];
    const bool amplicon = bool(std::stoi(argv[7]));
    const std::string file_out = argv[9], vcf_file = argv[10];
    const float auto_threshold = std::stof(argv[11]);
    { std::ofstream out(file_out); }
    {
        std::ofstream vcf(vcf_file);
        vcf << "##fileformat=VCFv4.2\n##source=call_variants\n"
               "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Total Depth\">\n"
               "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n";
    }
    omp_set_num_threads(std::max(1, num_threads));
    omp_set_max_active_levels(2);
    g_timing = std::getenv("HS_TIMING") != nullptr;
    phase("start");
    // CUDA context creation and module load take about a second: start them now, parse meanwhile
    int n_gpus = 1;
    if (const char* e = std::getenv("HSGPU_NGPUS")) n_gpus = std::max(1, std::atoi(e));
    int first_device = 0;
    if (const char* e = std::getenv("HSGPU_DEVICE")) first_device = std::atoi(e);
    std::vector<std::future<hsgpu_ctx*>> contexts;
    for (int g = 0; g < n_gpus; g++)
        contexts.push_back(std::async(std::getenv("HS_CTX_FIRST") ? std::launch::deferred : std::launch::async, [=]() {
            const double t0 = omp_get_wtime();
            hsgpu_ctx* ctx = nullptr;
            if (hsgpu_ctx_create(first_device + g, &ctx) != HSGPU_OK) return (hsgpu_ctx*)nullptr;
            if (g_timing) fprintf(stderr, "[hs timing] (context %d created in %.3f s on its own thread)\n", g, omp_get_wtime() - t0);
            return ctx;
        }));
    if (std::getenv("HS_CTX_FIRST"))  // experiment: create the contexts before anything else, one after the other
        for (auto& f : contexts) f.wait();
    Store st;
    std::cout << " - Loading all reads from " << reads_file << " in memory\n";
    parse_reads(reads_file, st);
    MappedFile reads_map;  // the sequences are packed straight from this mapping (process_batch)
    if (!reads_map.open(reads_file)) {
        std::cout << "problem reading files in index_reads, while trying to read " << reads_file << std::endl;
        std::exit(EXIT_FAILURE);
    }
    phase("parse_reads");
    std::cout << " - Loading all contigs from " << gfa_file << " in memory\n";
    parse_assembly(gfa_file, st);
    phase("parse_assembly");
    std::cout << " - Loading alignments of the reads on the contigs from " << sam_file << "\n";
    if (sam_file.size() >= 4 && sam_file.substr(sam_file.size() - 4, 4) == ".paf") {
        std::cout << "ERROR: please provide a .sam file as input for the alignments of the reads on the contigs." << std::endl;
        std::exit(EXIT_FAILURE);
    } else if (sam_file.size() >= 4 && sam_file.substr(sam_file.size() - 4, 4) == ".sam") {
        parse_sam(sam_file, st, amplicon);
    } else {
        std::cout << "ERROR: the file containing the alignments on the assembly should be .sam" << std::endl;
        std::exit(EXIT_FAILURE);
    }
    phase("parse_sam");
    std::cout << " - Calling variants on each contig\n";

    // the contigs to process (the reference skips one hard-coded debugging name, :1282)
    std::vector<int> todo;
    for (int ci = 0; ci < (int)st.contigs.size(); ci++)
        if (st.seqs[st.contigs[ci]].name != "edge_124@009") todo.push_back(ci);
    // shard over GPUs: heaviest contig first onto the least loaded device, then cut each device's list
    // into batches bounded by pileup cells
    std::vector<double> weight(st.contigs.size(), 0.0);
    for (int ci : todo) {
        double w = (double)st.seqs[st.contigs[ci]].sequence.size();
        for (int64_t id : st.seqs[st.contigs[ci]].alns) w += (double)(st.alns[id].pos_2_2 - st.alns[id].pos_2_1);
        weight[ci] = w;
    }
    std::vector<int> by_weight(todo);
    std::stable_sort(by_weight.begin(), by_weight.end(), [&](int a, int b) { return weight[a] > weight[b]; });
    std::vector<std::vector<int>> shard(n_gpus);
    std::vector<double> load(n_gpus, 0.0);
    for (int ci : by_weight) {
        const int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        shard[g].push_back(ci);
        load[g] += weight[ci];
    }
    const double batch_cells = 3e9;
    std::vector<ContigResult> results(st.contigs.size());
    std::vector<std::string> errors(n_gpus);
#pragma omp parallel for num_threads(n_gpus) schedule(static, 1)
    for (int g = 0; g < n_gpus; g++) {
        phase("shard setup");
        omp_set_num_threads(std::max(1, num_threads / n_gpus));  // host threads of this shard's inner loops
        hsgpu_ctx* ctx = nullptr;
        const std::function<hsgpu_ctx*()> get_ctx = [&]() {
            if (!ctx) {
                ctx = contexts[g].get();
                if (!ctx) {
                    std::cout << "ERROR: no usable GPU " << first_device + g << ": " << hsgpu_last_error(nullptr) << std::endl;
                    std::exit(1);  // there is no CPU fallback
                }
            }
            return ctx;
        };
        if (shard[g].empty()) {
            hsgpu_ctx_destroy(get_ctx());
            continue;
        }
        std::sort(shard[g].begin(), shard[g].end());
        std::vector<int> batch;
        double cells = 0;
        for (size_t i = 0; i < shard[g].size(); i++) {
            batch.push_back(shard[g][i]);
            cells += weight[shard[g][i]];
            if (cells >= batch_cells || i + 1 == shard[g].size()) {
                process_batch(get_ctx, st, reads_map, batch, auto_threshold, results);
                batch.clear();
                cells = 0;
            }
        }
        hsgpu_ctx_destroy(get_ctx());
    }

    phase("gpu shards total");
    // same accumulation order as the reference at one thread: contig by contig
    float total_error_rate = 0;
    int n_rated = 0;
    std::unordered_map<int, std::vector<Column>> variants;
    for (int ci : todo) {
        if (results[ci].mean_distance > 0) {
            total_error_rate += results[ci].mean_distance;
            n_rated += 1;
        }
        st.seqs[st.contigs[ci]].depth = results[ci].depth;
        variants[(int)st.contigs[ci]] = std::move(results[ci].merged);
    }
    std::ofstream error_file(error_rate_out);
    std::cout << "total error rate : " << total_error_rate << " number of contigs : " << n_rated << std::endl;
    error_file << total_error_rate / n_rated << std::endl;
    error_file.close();
    write_outputs(st, variants, file_out, vcf_file);
    phase("write_outputs");
    if (g_timing) fprintf(stderr, "[hs timing] %-28s %8.3f s\n", "main() so far", omp_get_wtime() - t_main);
    // Every output file is written and closed. Leaving through exit() would now spend a few hundred milliseconds
    // unwinding: the CUDA runtime's atexit handler tears the primary context down and the parsed reads (hundreds of
    // megabytes of strings) are freed one by one. The operating system reclaims both at once.
    std::cout.flush();
    std::cerr.flush();
    fflush(nullptr);
    if (!std::getenv("HS_FULL_TEARDOWN")) _exit(0);
    return 0;
}
