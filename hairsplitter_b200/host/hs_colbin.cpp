// Binary sidecar of the .col file; see hs_colbin.h for what it is and when it is used.
//
// File layout (little-endian, every section starts on an 8-byte boundary):
//   Header    magic "HSCOLB01", col_size, col_mtime_ns, col_sig, n_contigs
//   Directory n_contigs x {offset, bytes} of the contig blocks, in the order of the .col file
//   Block     {head_bytes, n_snps, n_cells}, head text, int32 pos[n_snps], u8 ref[n_snps], u8 second[n_snps],
//             u64 off[n_snps + 1], u32 read_idx[n_cells], u8 code[n_cells]
#include "hs_colbin.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace hs {

namespace {

const char kMagic[8] = {'H', 'S', 'C', 'O', 'L', 'B', '0', '1'};
const size_t kSigSpan = 64 * 1024;

struct Header {
    char magic[8];
    uint64_t col_size;
    uint64_t col_mtime_ns;
    uint64_t col_sig;
    uint64_t n_contigs;
};
struct DirEntry {
    uint64_t off, bytes;
};
struct BlockHeader {
    uint64_t head_bytes, n_snps, n_cells;
};

inline size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }

uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) {
        h ^= b[i];
        h *= 1099511628211ull;
    }
    return h;
}

// identity of a .col file: size, mtime and a hash over its first and last 64 KiB
bool col_identity(const std::string& col_path, uint64_t& size, uint64_t& mtime_ns, uint64_t& sig) {
    struct stat sb;
    // regular files only: opening a pipe for reading would wait for a writer that never comes
    if (::stat(col_path.c_str(), &sb) != 0 || !S_ISREG(sb.st_mode)) return false;
    const int fd = ::open(col_path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    if (::fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
        ::close(fd);
        return false;
    }
    size = (uint64_t)sb.st_size;
    mtime_ns = (uint64_t)sb.st_mtim.tv_sec * 1000000000ull + (uint64_t)sb.st_mtim.tv_nsec;
    std::vector<char> buf(kSigSpan);
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, &size, sizeof(size));
    const size_t n_head = (size_t)std::min<uint64_t>(size, kSigSpan);
    bool ok = ::pread(fd, buf.data(), n_head, 0) == (ssize_t)n_head;
    h = fnv1a(h, buf.data(), n_head);
    if (ok && size > kSigSpan) {
        const size_t n_tail = (size_t)std::min<uint64_t>(size - kSigSpan, kSigSpan);
        ok = ::pread(fd, buf.data(), n_tail, (off_t)(size - n_tail)) == (ssize_t)n_tail;
        h = fnv1a(h, buf.data(), n_tail);
    }
    ::close(fd);
    sig = h;
    return ok;
}

size_t block_bytes(size_t head, size_t n_snps, size_t n_cells) {
    return sizeof(BlockHeader) + pad8(head) + pad8(4 * n_snps) + pad8(n_snps) + pad8(n_snps) + 8 * (n_snps + 1) +
           pad8(4 * n_cells) + pad8(n_cells);
}

}  // namespace

bool col_sidecar_enabled() {
    const char* e = std::getenv("HS_SIDECAR");
    return !(e && e[0] == '0' && e[1] == 0);
}

std::string col_sidecar_path(const std::string& col_path) { return col_path + ".hsb"; }

bool write_col_sidecar(const std::string& col_path, const std::vector<ColSidecarBlock>& blocks) {
    const std::string path = col_sidecar_path(col_path);
    ::unlink(path.c_str());  // whatever happens below, no stale sidecar stays next to a new .col
    if (!col_sidecar_enabled()) return false;
    Header h;
    std::memcpy(h.magic, kMagic, 8);
    h.n_contigs = blocks.size();
    if (!col_identity(col_path, h.col_size, h.col_mtime_ns, h.col_sig)) return false;
    // sizes first, so that every block can be laid out by its own thread
    const size_t n = blocks.size();
    std::vector<DirEntry> dir(n);
    std::vector<size_t> n_cells(n, 0);
    size_t at = pad8(sizeof(Header)) + n * sizeof(DirEntry);
    for (size_t b = 0; b < n; b++) {
        size_t cells = 0;
        for (const Column& c : *blocks[b].snps) cells += std::min(c.readIdxs.size(), c.content.size());
        n_cells[b] = cells;
        dir[b].off = at;
        dir[b].bytes = block_bytes(blocks[b].head.size(), blocks[b].snps->size(), cells);
        at += dir[b].bytes;
    }
    const size_t total = at;
    const std::string tmp = path + ".tmp";
    const int fd = ::open(tmp.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return false;
    if (::ftruncate(fd, (off_t)total) != 0) {
        ::close(fd);
        ::unlink(tmp.c_str());
        return false;
    }
    char* base = (char*)::mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    if (base == (char*)MAP_FAILED) {
        ::unlink(tmp.c_str());
        return false;
    }
    std::memcpy(base, &h, sizeof(h));
    if (n) std::memcpy(base + pad8(sizeof(Header)), dir.data(), n * sizeof(DirEntry));
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t b = 0; b < n; b++) {
        const std::vector<Column>& snps = *blocks[b].snps;
        const size_t S = snps.size(), head = blocks[b].head.size();
        char* p = base + dir[b].off;
        BlockHeader bh{head, S, n_cells[b]};
        std::memcpy(p, &bh, sizeof(bh));
        p += sizeof(bh);
        std::memcpy(p, blocks[b].head.data(), head);
        p += pad8(head);
        int32_t* pos = (int32_t*)p;
        p += pad8(4 * S);
        uint8_t* ref = (uint8_t*)p;
        p += pad8(S);
        uint8_t* second = (uint8_t*)p;
        p += pad8(S);
        uint64_t* off = (uint64_t*)p;
        p += 8 * (S + 1);
        uint32_t* idx = (uint32_t*)p;
        p += pad8(4 * n_cells[b]);
        uint8_t* code = (uint8_t*)p;
        uint64_t o = 0;
        for (size_t s = 0; s < S; s++) {
            const Column& c = snps[s];
            // the text holds every read index but only min(indices, codes) codes, and the parser walks the codes
            const size_t m = std::min(c.readIdxs.size(), c.content.size());
            pos[s] = c.pos;
            ref[s] = c.ref_base;
            second[s] = c.second_base;
            off[s] = o;
            if (m) {
                std::memcpy(idx + o, c.readIdxs.data(), 4 * m);
                std::memcpy(code + o, c.content.data(), m);
            }
            o += m;
        }
        off[S] = o;
    }
    const bool synced = ::msync(base, total, MS_ASYNC) == 0;
    ::munmap(base, total);
    if (!synced || ::rename(tmp.c_str(), path.c_str()) != 0) {
        ::unlink(tmp.c_str());
        return false;
    }
    return true;
}

bool read_col_sidecar(const std::string& col_path, std::vector<ColContig>& contigs, int max_coverage,
                      float rarest_strain_abundance) {
    if (!col_sidecar_enabled()) return false;
    const std::string path = col_sidecar_path(col_path);
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    if (::fstat(fd, &sb) != 0 || (size_t)sb.st_size < sizeof(Header)) {
        ::close(fd);
        return false;
    }
    const size_t size = (size_t)sb.st_size;
    const char* base = (const char*)::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (base == (const char*)MAP_FAILED) return false;
    struct Unmap {
        const char* p;
        size_t n;
        ~Unmap() { ::munmap((void*)p, n); }
    } unmap{base, size};
    Header h;
    std::memcpy(&h, base, sizeof(h));
    if (std::memcmp(h.magic, kMagic, 8) != 0) return false;
    uint64_t col_size, col_mtime, col_sig;
    if (!col_identity(col_path, col_size, col_mtime, col_sig)) return false;
    if (col_size != h.col_size || col_mtime != h.col_mtime_ns || col_sig != h.col_sig) return false;
    const size_t n = (size_t)h.n_contigs;
    const size_t dir_at = pad8(sizeof(Header));
    if (n > (size - dir_at) / sizeof(DirEntry)) return false;
    const DirEntry* dir = (const DirEntry*)(base + dir_at);
    // every block must lie inside the file and be exactly as large as its own counts say
    for (size_t b = 0; b < n; b++) {
        if (dir[b].off % 8 || dir[b].off > size || dir[b].bytes > size - dir[b].off || dir[b].bytes < sizeof(BlockHeader))
            return false;
        BlockHeader bh;
        std::memcpy(&bh, base + dir[b].off, sizeof(bh));
        if (bh.head_bytes > dir[b].bytes || bh.n_snps > dir[b].bytes || bh.n_cells > dir[b].bytes) return false;
        if (block_bytes(bh.head_bytes, bh.n_snps, bh.n_cells) != dir[b].bytes) return false;
    }
    ::madvise((void*)base, size, MADV_SEQUENTIAL);
    std::vector<ColContig> out(n);
    bool bad = false;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t b = 0; b < n; b++) {
        const char* p = base + dir[b].off;
        BlockHeader bh;
        std::memcpy(&bh, p, sizeof(bh));
        p += sizeof(bh);
        const size_t S = (size_t)bh.n_snps, cells = (size_t)bh.n_cells;
        ColContig& c = out[b];
        parse_col_head(p, p + bh.head_bytes, c);
        p += pad8(bh.head_bytes);
        const int32_t* pos = (const int32_t*)p;
        p += pad8(4 * S);
        const uint8_t* ref = (const uint8_t*)p;
        p += pad8(S);
        const uint8_t* second = (const uint8_t*)p;
        p += pad8(S);
        const uint64_t* off = (const uint64_t*)p;
        p += 8 * (S + 1);
        const uint32_t* idx = (const uint32_t*)p;
        p += pad8(4 * cells);
        const uint8_t* code = (const uint8_t*)p;
        if (off[0] != 0 || off[S] != cells) {
#pragma omp atomic write
            bad = true;
            continue;
        }
        c.snps.reserve(S);
        for (size_t s = 0; s < S; s++) {
            if (off[s] > off[s + 1] || off[s + 1] > cells) {
#pragma omp atomic write
                bad = true;
                break;
            }
            const size_t m = (size_t)(off[s + 1] - off[s]);
            const uint32_t* si = idx + off[s];
            const uint8_t* sc = code + off[s];
            const uint8_t ref_base = ref[s], second_base = second[s];
            Column snp;
            snp.pos = pos[s];
            snp.ref_base = ref_base;
            snp.second_base = second_base;
            snp.content.reserve(m);
            snp.readIdxs.reserve(m);
            // the loop of parse_column_file over the tokens of the line (:150-166), on the arrays
            int cov_maj = 0, cov_sec = 0, cov = 0;
            for (size_t k = 0; k < m; k++) {
                const int r = (int)si[k];
                if (sc[k] != ' ' && cov < max_coverage) {
                    snp.content.push_back(sc[k]);
                    snp.readIdxs.push_back((uint32_t)r);
                    if (sc[k] == ref_base) cov_maj++;
                    else if (sc[k] == second_base) cov_sec++;
                }
                if (sc[k] != ' ' && r >= 0) cov++;
            }
            if ((float)cov_sec >= rarest_strain_abundance * (float)(cov_maj + cov_sec)) c.snps.push_back(std::move(snp));
        }
    }
    if (bad) return false;
    contigs.reserve(contigs.size() + n);
    for (ColContig& c : out) contigs.push_back(std::move(c));
    return true;
}

}  // namespace hs
