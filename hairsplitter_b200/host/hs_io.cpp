#include "hs_io.h"

#include "hs_colbin.h"

#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <cstring>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>

namespace hs {

using std::string;

static string up_to_first_blank(const string& s) {
    const size_t p = s.find(' ');
    return p == string::npos ? s : s.substr(0, p);
}

// src/input_output.cpp:39-110. Records are delimited by header lines ('>' for *.fa / *.fasta, '@'
// otherwise; a '@' line only opens a FASTQ record when the record in progress is complete or the
// previous line was not the '+' separator). A record keeps its name (header up to the first blank), the
// length of its FIRST sequence line and the file offset of that line; the sequence itself is read
// back on demand.
void parse_reads(const string& path, Store& st) {
    char format = '@';
    if ((path.size() > 6 && path.substr(path.size() - 6, 6) == ".fasta") || path.substr(path.size() - 3, 3) == ".fa")
        format = '>';
    std::ifstream in(path);
    if (!in) {
        std::cout << "problem reading files in index_reads, while trying to read " << path << std::endl;
        throw std::invalid_argument("Input file could not be read");
    }
    std::vector<string> record;
    string line;
    int64_t record_start = 0, offset = 0;
    char previous_first = '+';
    auto flush = [&]() {
        SeqRec r;
        r.name = up_to_first_blank(record.at(0).substr(1));
        r.length = (int64_t)record.at(1).size();
        r.file_pos = record_start + (int64_t)record[0].size() + 1;
        st.index[r.name] = (int64_t)st.seqs.size();
        st.seqs.push_back(std::move(r));
    };
    while (std::getline(in, line)) {
        const bool header = !line.empty() && line[0] == format && record.size() >= 2 &&
                            (format == '>' || previous_first != '+' || record.size() == 4);
        if (header) {
            flush();
            record_start = offset;
            record.clear();
        }
        record.push_back(line);
        if (!line.empty()) previous_first = line[0];
        offset += 1 + (int64_t)line.size();
    }
    flush();
}

// src/input_output.cpp:120-263: S lines give contigs (name up to the first blank, sequence, optional
// dp:/DP: depth tag); L lines must name known segments.
void parse_assembly(const string& path, Store& st) {
    std::ifstream in(path);
    if (!in) {
        std::cout << "problem reading files in index_reads, while trying to read " << path << std::endl;
        throw std::invalid_argument("Input file could not be read");
    }
    string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] == 'S') {
            std::istringstream fields(line);
            string field, name;
            int k = 0;
            while (std::getline(fields, field, '\t')) {
                if (k == 1) {
                    name = up_to_first_blank(field);
                } else if (k == 2) {
                    SeqRec r;
                    r.name = name;
                    r.length = (int64_t)field.size();
                    r.sequence = field;
                    st.index[name] = (int64_t)st.seqs.size();
                    st.contigs.push_back((int64_t)st.seqs.size());
                    st.seqs.push_back(std::move(r));
                } else if (field.substr(0, 2) == "dp" || field.substr(0, 2) == "DP") {
                    st.seqs[st.contigs.back()].depth = (float)std::atoi(field.substr(5, field.size() - 5).c_str());
                }
                k++;
            }
        } else if (line[0] == 'L') {
            std::istringstream fields(line);
            string field, n1, n2;
            int k = 0;
            bool ok = true;
            while (std::getline(fields, field, '\t')) {
                if (k == 1) n1 = field;
                else if (k == 3) n2 = field;
                else if (k == 2 || k == 4) {
                    const string& nm = (k == 2) ? n1 : n2;
                    if ((field != "+" && field != "-") || st.index.find(nm) == st.index.end()) ok = false;
                }
                k++;
            }
            if (!ok) {
                std::cout << "Problem in reading the link : " << line << std::endl;
                std::cout << "Problem while reading GFA file " + path +
                                 ". Ensure that all the contigs described in 'L' lines are present in 'S' lines."
                          << std::endl;
                throw std::invalid_argument("Invalid GFA");
            }
        }
    }
}

void cigar_ops(const string& cigar, std::vector<uint32_t>& ops) {
    ops.clear();
    if (cigar == "*") return;
    // index in "MIDNSHP=X"; a letter no loop of the reference looks at behaves like padding (6)
    static const struct OpCodes {
        uint8_t code[256];
        OpCodes() {
            for (int c = 0; c < 256; c++) code[c] = 6;
            const char* letters = "MIDNSHP=X";
            for (int k = 0; letters[k]; k++) code[(unsigned char)letters[k]] = (uint8_t)k;
        }
    } table;
    const char* const begin = cigar.data();
    const char* const end = begin + cigar.size();
    uint32_t n = 0;
    const char* num = begin;  // first digit of the length being read
    for (const char* p = begin; p < end; p++) {
        const unsigned d = (unsigned)(*p - '0');
        if (d <= 9) {
            n = n * 10 + d;
            continue;
        }
        const size_t n_digits = (size_t)(p - num);
        if (n_digits == 0 || n_digits > 9) {  // no length, or one that may not fit: what std::stoi makes of it
            try {
                n = (uint32_t)std::stoi(string(num, n_digits));
            } catch (...) {
                std::cout << "ERROR : could not convert " << cigar << " to int" << std::endl;
                std::exit(1);
            }
        }
        if ((int)n > 0) ops.push_back(((uint32_t)n << 4) | (uint32_t)table.code[(unsigned char)*p]);
        n = 0;
        num = p + 1;
    }
}

// length of the run of `letter` ops at the start / end of the CIGAR, the way parse_SAM measures clips
// (src/input_output.cpp:387-470): only the first op counts at the start; at the end everything behind
// the last letter that is not `letter` is taken as one number.
static int clip_at_start(const string& cigar, char letter) {
    string digits;
    for (size_t i = 0; i < cigar.size(); i++) {
        if (cigar[i] > '9' || cigar[i] < '0') {
            if (cigar[i] != letter) digits.clear();
            break;
        }
        digits += cigar[i];
    }
    return digits.empty() ? 0 : std::stoi(digits);
}
static int clip_at_end(const string& cigar, char letter) {
    string tail;
    for (int i = (int)cigar.size() - 1; i >= 0; i--) {
        if ((cigar[i] > '9' || cigar[i] < '0') && cigar[i] != letter) break;
        tail = cigar[i] + tail;
    }
    if (tail.empty()) return 0;
    tail = tail.substr(0, tail.size() - 1);
    return std::stoi(tail);
}

// src/input_output.cpp:274-536. Everything that depends on one line only (field splitting, flag tests, clip
// lengths, aligned lengths) is done for all lines in parallel; the name lookups that can CHANGE the index
// (the reference's map access registers unknown names with index 0, so a second record of an unknown read
// lands on read 0) and the bookkeeping are replayed in file order afterwards.
namespace {
struct SamRecord {
    std::string read_name, contig_name, cigar;
    int64_t read = -1, contig = -1;  // -1: name not in the index when the line was parsed
    int pos_1_1 = 0, pos_1_2 = 0, pos_2_1 = 0, pos_2_2 = 0;
    bool forward = true, good = true, enough_fields = false;
    int n_fields = 0;
};

void parse_sam_line(const char* p, const char* end, const Store& st, bool amplicon, SamRecord& rec) {
    std::string field, cigar;
    std::vector<uint32_t> ops;
    int read_length = 0, pos = -1, flag = 0, nonmatching = 0, k = 0;
    while (p <= end) {  // std::getline semantics: a trailing tab does not open an empty last field
        const char* q = (const char*)memchr(p, '\t', (size_t)(end - p));
        if (!q) q = end;
        if (q == end && p == end) break;
        field.assign(p, q);
        if (k == 0) {
            rec.read_name = field;
            auto it = st.index.find(field);
            if (it != st.index.end()) rec.read = it->second;
        } else if (k == 1) {
            flag = std::stoi(field);
            if (flag % 8 >= 4) rec.good = false;
            if (flag % 32 >= 16) rec.forward = false;
        } else if (k == 2) {
            rec.contig_name = field;
            auto it = st.index.find(field);
            if (it != st.index.end()) rec.contig = it->second;
        } else if (k == 3) {
            pos = std::stoi(field);
        } else if (k == 5) {
            cigar = field;
        } else if (field.compare(0, 5, "LN:i:") == 0) {
            read_length = std::stoi(field.substr(5, field.size() - 5));
        } else if (field.compare(0, 5, "NM:i:") == 0) {
            nonmatching = std::stoi(field.substr(5, field.size() - 5));
        }
        k++;
        p = q + 1;
    }
    rec.enough_fields = k > 10;
    rec.n_fields = k;
    if (!rec.good || !rec.enough_fields) return;
    cigar_ops(cigar, ops);
    int h_start = clip_at_start(cigar, 'H'), h_end = clip_at_end(cigar, 'H');
    if (!rec.forward) std::swap(h_start, h_end);
    int s_start = clip_at_start(cigar, 'S'), s_end = clip_at_end(cigar, 'S');
    if (!rec.forward) std::swap(s_start, s_end);
    if (h_start + h_end > 0.2 * read_length && flag < 2048) rec.good = false;
    else if (flag % 512 >= 256) rec.good = false;
    if (amplicon && nonmatching > 0.2 * read_length) rec.good = false;
    if (!rec.good) return;
    int on_read = 0, on_contig = 0;
    for (uint32_t op : ops) {
        const int n = (int)(op >> 4), ty = (int)(op & 15);
        if (ty == 0 || ty == 7 || ty == 8) { on_read += n; on_contig += n; }
        else if (ty == 1) on_read += n;
        else if (ty == 2) on_contig += n;
    }
    rec.pos_1_1 = s_start + h_start;
    rec.pos_1_2 = s_start + h_start + on_read;
    rec.pos_2_1 = pos - 1;
    rec.pos_2_2 = pos + on_contig;
    rec.cigar.swap(cigar);
}
}  // namespace

bool MappedFile::open(const string& path) {
    close();
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    if (::fstat(fd, &sb) != 0) {
        ::close(fd);
        return false;
    }
    size = (size_t)sb.st_size;
    if (size > 0) {
        void* m = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) {
            ::close(fd);
            size = 0;
            return false;
        }
        ::madvise(m, size, MADV_WILLNEED);
        data = (const char*)m;
    }
    ::close(fd);
    return true;
}

void MappedFile::close() {
    if (data) ::munmap((void*)data, size);
    data = nullptr;
    size = 0;
}

void parse_sam(const string& path, Store& st, bool amplicon) {
    MappedFile text;
    if (!text.open(path)) {
        std::cout << "problem reading SAM file " << path << std::endl;
        throw std::invalid_argument("Input file '" + path + "' could not be read");
    }
    // line starts: found in parallel over slices of the mapping, concatenated in file order
    std::vector<size_t> starts;
    {
        const int nt = std::max(1, omp_get_max_threads());
        std::vector<std::vector<size_t>> found((size_t)nt);
#pragma omp parallel for schedule(static, 1)
        for (int t = 0; t < nt; t++) {
            size_t lo = text.size * (size_t)t / (size_t)nt, hi = text.size * (size_t)(t + 1) / (size_t)nt;
            if (t > 0 && lo > 0) {  // first line that starts inside the slice
                const void* nl = memchr(text.data + lo - 1, '\n', text.size - (lo - 1));
                lo = nl ? (size_t)((const char*)nl - text.data) + 1 : text.size;
            }
            for (size_t p = lo; p < hi;) {
                found[t].push_back(p);
                const void* nl = memchr(text.data + p, '\n', text.size - p);
                p = nl ? (size_t)((const char*)nl - text.data) + 1 : text.size;
            }
        }
        size_t total = 0;
        for (const auto& v : found) total += v.size();
        starts.reserve(total + 1);
        for (const auto& v : found) starts.insert(starts.end(), v.begin(), v.end());
    }
    const size_t n_lines = starts.size();
    starts.push_back(text.size + 1);  // as if the last line ended with a newline
    std::vector<SamRecord> recs(n_lines);
    std::vector<char> is_record(n_lines, 0);
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t i = 0; i < n_lines; i++) {
        const char* p = text.data + starts[i];
        const char* end = text.data + std::min(starts[i + 1] - 1, text.size);  // the newline (or the end)
        if (p < end && *p == '@') continue;
        is_record[i] = 1;
        parse_sam_line(p, end, st, amplicon, recs[i]);
    }
    for (size_t i = 0; i < n_lines; i++) {
        if (!is_record[i]) continue;
        SamRecord& rec = recs[i];
        bool good = rec.good;
        if (rec.n_fields == 0) continue;  // an empty line: the reference looks nothing up
        int64_t read = rec.read, contig = rec.contig;
        if (read < 0) {
            if (st.index.find(rec.read_name) == st.index.end()) {
                std::cout << "WARNING: read in the sam file not found in reads file, ignoring: " << rec.read_name << std::endl;
                good = false;
            }
            read = st.index[rec.read_name];  // registers the name (index 0), like the reference's map access
        }
        if (contig < 0 && rec.n_fields >= 3) contig = st.index[rec.contig_name];
        if (!(good && rec.enough_fields && contig != read)) continue;
        Alignment a;
        a.read = read;
        a.contig = contig;
        a.pos_1_1 = rec.pos_1_1;
        a.pos_1_2 = rec.pos_1_2;
        a.pos_2_1 = rec.pos_2_1;
        a.pos_2_2 = rec.pos_2_2;
        a.strand = rec.forward;
        a.cigar.swap(rec.cigar);
        const int64_t id = (int64_t)st.alns.size();
        st.seqs[read].alns.push_back(id);
        st.seqs[contig].alns.push_back(id);
        st.alns.push_back(std::move(a));
    }
}

void load_read_sequences(std::ifstream& reads_file, const Store& st, int64_t contig, std::vector<string>& out) {
    out.clear();
    string line;
    for (int64_t id : st.seqs[contig].alns) {
        const Alignment& a = st.alns[id];
        const int64_t read = (a.read != contig) ? a.read : a.contig;
        reads_file.clear();
        reads_file.seekg(st.seqs[read].file_pos);
        std::getline(reads_file, line);
        out.push_back(line);
    }
}

void view_read_sequences(const MappedFile& reads, const Store& st, int64_t contig,
                         std::vector<std::pair<const char*, size_t>>& out) {
    out.clear();
    for (int64_t id : st.seqs[contig].alns) {
        const Alignment& a = st.alns[id];
        const int64_t read = (a.read != contig) ? a.read : a.contig;
        // what seekg(file_pos) + getline returns: the bytes up to the next newline (or the end of the file)
        const size_t at = (size_t)st.seqs[read].file_pos;
        size_t len = 0;
        if (at < reads.size) {
            const void* nl = memchr(reads.data + at, '\n', reads.size - at);
            len = nl ? (size_t)((const char*)nl - (reads.data + at)) : reads.size - at;
        }
        out.emplace_back(reads.data + std::min(at, reads.size), len);
    }
}

namespace {
struct BaseCodes {
    uint8_t code[256];
    BaseCodes() {
        for (int c = 0; c < 256; c++) code[c] = 3;
        code[(unsigned char)'A'] = 0;
        code[(unsigned char)'C'] = 1;
        code[(unsigned char)'G'] = 2;
    }
};
const BaseCodes kBaseCodes;
}  // namespace

void pack_bases_2bit(const char* seq, int64_t n, uint32_t* out) {
    const uint8_t* const lut = kBaseCodes.code;
    const unsigned char* s = (const unsigned char*)seq;
    const int64_t full = n / 16;
    for (int64_t w = 0; w < full; w++, s += 16) {
        // four independent chains of four bases
        const uint32_t a = lut[s[0]] | (lut[s[1]] << 2) | (lut[s[2]] << 4) | (lut[s[3]] << 6);
        const uint32_t b = lut[s[4]] | (lut[s[5]] << 2) | (lut[s[6]] << 4) | (lut[s[7]] << 6);
        const uint32_t c = lut[s[8]] | (lut[s[9]] << 2) | (lut[s[10]] << 4) | (lut[s[11]] << 6);
        const uint32_t d = lut[s[12]] | (lut[s[13]] << 2) | (lut[s[14]] << 4) | (lut[s[15]] << 6);
        out[w] = a | (b << 8) | (c << 16) | (d << 24);
    }
    const int rest = (int)(n - 16 * full);
    if (rest > 0) {
        uint32_t v = 0;
        for (int j = 0; j < rest; j++) v |= (uint32_t)lut[s[j]] << (2 * j);
        out[full] = v;
    }
}

// src/call_variants.cpp:1174-1213. Alleles are written as decimal integers, lists end with a comma, every
// contig block ends with an empty line; the .vcf is reopened, which drops the header written earlier.
// The text of every contig block is formatted in parallel, then written in the iteration order of the container, like
// the reference's sequential loop. The SNPS lines are the bulk (two numbers per cell, 1.4 GB on BASELINE config 3): their
// exact size is counted first, every block is formatted once into a buffer of that size, and the blocks go to the file
// at their offsets from all threads (pwrite) instead of through one stream.
static inline int uint_digits(unsigned v) {
    return 1 + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u) + (v >= 100000u) + (v >= 1000000u) + (v >= 10000000u) +
           (v >= 100000000u) + (v >= 1000000000u);
}
static inline char* put_uint(char* p, unsigned v) {
    const int n = uint_digits(v);
    for (int k = n - 1; k >= 0; k--) {
        p[k] = (char)('0' + v % 10u);
        v /= 10u;
    }
    return p + n;
}
static inline char* put_text(char* p, const string& t) {
    std::memcpy(p, t.data(), t.size());
    return p + t.size();
}
static bool write_all_at(int fd, const char* data, size_t n, off_t at) {
    while (n > 0) {
        const ssize_t w = ::pwrite(fd, data, n, at);
        if (w <= 0) return false;
        data += w;
        n -= (size_t)w;
        at += w;
    }
    return true;
}

void write_outputs(const Store& st, const std::unordered_map<int, std::vector<Column>>& variants,
                   const string& col_file, const string& vcf_file) {
    std::vector<const std::pair<const int, std::vector<Column>>*> order;
    for (const auto& kv : variants) order.push_back(&kv);
    const size_t nb = order.size();
    const bool timing = std::getenv("HS_TIMING") != nullptr;
    double t_phase = omp_get_wtime();
    auto phase = [&](const char* what) {
        if (!timing) return;
        const double t = omp_get_wtime();
        fprintf(stderr, "[hs timing]   write_outputs: %-18s %8.3f s\n", what, t - t_phase);
        t_phase = t;
    };
    std::vector<string> head_text(nb), vcf_text(nb);  // CONTIG + READ lines of the block (the sidecar keeps them as text)
    std::vector<std::unique_ptr<char[]>> snp_text(nb);
    std::vector<size_t> snp_bytes(nb, 0);
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t i = 0; i < nb; i++) {
        const SeqRec& contig = st.seqs[order[i]->first];
        string& o = head_text[i];
        string& v = vcf_text[i];
        {
            std::ostringstream head;  // the depth is a float printed with the stream's default formatting
            head << "CONTIG\t" << contig.name << "\t" << contig.sequence.size() << "\t" << contig.depth << "\n";
            o += head.str();
        }
        for (int64_t id : contig.alns) {
            const Alignment& a = st.alns[id];
            o += "READ\t";
            o += st.seqs[a.read].name;
            for (const int x : {a.pos_1_1, a.pos_1_2, a.pos_2_1, a.pos_2_2}) {
                o += '\t';
                o += std::to_string(x);
            }
            o += a.strand ? "\t1\n" : "\t0\n";
        }
        const std::vector<Column>& snps = order[i]->second;
        std::vector<string> pos_text(snps.size());  // a signed int: printed by the library
        size_t bytes = 1;                            // the empty line that ends the block
        for (size_t k = 0; k < snps.size(); k++) {
            const Column& c = snps[k];
            pos_text[k] = std::to_string(c.pos);
            const size_t n_code = std::min(c.content.size(), c.readIdxs.size());
            bytes += 5 + pos_text[k].size() + 1 + (size_t)uint_digits(c.ref_base) + 1 + (size_t)uint_digits(c.second_base) + 1 +
                     c.readIdxs.size() + 1 + n_code + 1;
            for (const uint32_t r : c.readIdxs) bytes += (size_t)uint_digits(r);
            for (size_t r = 0; r < n_code; r++) bytes += (size_t)uint_digits(c.content[r]);
        }
        snp_bytes[i] = bytes;
        snp_text[i].reset(new char[bytes]);
        char* p = snp_text[i].get();
        for (size_t k = 0; k < snps.size(); k++) {
            const Column& c = snps[k];
            std::memcpy(p, "SNPS\t", 5);
            p = put_text(p + 5, pos_text[k]);
            *p++ = '\t';
            p = put_uint(p, (unsigned)(int)c.ref_base);
            *p++ = '\t';
            p = put_uint(p, (unsigned)(int)c.second_base);
            *p++ = '\t';
            for (const uint32_t r : c.readIdxs) {
                p = put_uint(p, (unsigned)r);
                *p++ = ',';
            }
            *p++ = '\t';
            for (size_t r = 0; r < c.content.size() && r < c.readIdxs.size(); r++) {
                p = put_uint(p, (unsigned)(int)c.content[r]);
                *p++ = ',';
            }
            *p++ = '\n';
            v += contig.name;
            v += '\t';
            v += pos_text[k];
            v += "\t.\t";
            v += "ACGT-"[(c.ref_base - '!') % 5];
            v += '\t';
            v += "ACGT-"[(c.second_base - '!') % 5];
            v += "\t.\t.\tDP=";
            v += std::to_string(c.readIdxs.size());
            v += '\n';
        }
        *p++ = '\n';
        if ((size_t)(p - snp_text[i].get()) != bytes) {  // the count above and the formatting must agree
            std::cout << "ERROR: internal: .col block size mismatch" << std::endl;
            std::exit(1);
        }
        v += '\n';
    }
    phase("format");
    // the .col file: every block at its offset
    std::vector<size_t> at(nb + 1, 0);
    for (size_t i = 0; i < nb; i++) at[i + 1] = at[i] + head_text[i].size() + snp_bytes[i];
    {
        const int fd = ::open(col_file.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        bool ok = fd >= 0;
        if (ok && ::ftruncate(fd, (off_t)at[nb]) != 0) {
            // not a regular file (a pipe, a device): one block after the other, like the reference's stream
            for (size_t i = 0; i < nb && ok; i++) {
                const char* const piece[2] = {head_text[i].data(), snp_text[i].get()};
                const size_t piece_bytes[2] = {head_text[i].size(), snp_bytes[i]};
                for (int k = 0; k < 2; k++) {
                    const char* data = piece[k];
                    size_t n = piece_bytes[k];
                    while (n > 0 && ok) {
                        const ssize_t w = ::write(fd, data, n);
                        if (w <= 0) ok = false;
                        else { data += w; n -= (size_t)w; }
                    }
                }
            }
        } else if (ok) {
#pragma omp parallel for schedule(dynamic, 1)
            for (size_t i = 0; i < nb; i++) {
                const bool w = write_all_at(fd, head_text[i].data(), head_text[i].size(), (off_t)at[i]) &&
                               write_all_at(fd, snp_text[i].get(), snp_bytes[i], (off_t)(at[i] + head_text[i].size()));
                if (!w) {
#pragma omp atomic write
                    ok = false;
                }
            }
        }
        if (fd >= 0) ::close(fd);
        if (!ok) {
            std::cout << "ERROR: cannot write " << col_file << std::endl;
            std::exit(1);
        }
    }
    phase(".col");
    {
        std::ofstream vcf(vcf_file, std::ios::binary);
        for (size_t i = 0; i < nb; i++) vcf.write(vcf_text[i].data(), (std::streamsize)vcf_text[i].size());
    }
    phase(".vcf");
    // the same content once more as flat arrays, for HS_separate_reads (hs_colbin.h)
    std::vector<ColSidecarBlock> blocks(nb);
    for (size_t i = 0; i < nb; i++) {
        blocks[i].head = std::move(head_text[i]);
        blocks[i].snps = &order[i]->second;
    }
    write_col_sidecar(col_file, blocks);
    phase("sidecar");
}

}  // namespace hs
