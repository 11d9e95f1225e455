// TEST INFRASTRUCTURE ONLY (oracle/): runs the UNMODIFIED read-clipping loop body of the reference's modify_GFA
// (src/create_new_contigs.cpp:392-447) on one (CIGAR, interval) pair. The loop body is not a function in the
// reference, so the build recipe (oracle/Makefile, target _ref/clip_snippet.inc) cuts those lines out of the source
// where it lies under /root/reference and this file includes them between declarations of the names they use.
// No reference source is stored in this repository; the snippet exists only under oracle/_ref/ (git-ignored).
#include <algorithm>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

std::string convert_cigar(std::string& cigar);   // src/tools.cpp:27 (compiled reference object)
std::string convert_cigar2(std::string& cigar);  // src/tools.cpp:61

using std::max;
using std::min;
using std::string;

namespace {
struct ShimOverlap {
    int position_2_1;
    std::string CIGAR;
};
struct ShimRead {
    std::vector<int> neighbors_;
};
}  // namespace

extern "C" {

// out = {posOnReadStart, posOnReadEnd, posOnCIGARStart, posOnCIGAREnd}; returns 0 or -2 (interval.second[r] = -2)
int hsref_clip_read(const char* cigar, int pos_2_1, int leftToPolish, int rightToPolish, int* out) {
    std::vector<ShimOverlap> allOverlaps(1);
    allOverlaps[0].position_2_1 = pos_2_1;
    allOverlaps[0].CIGAR = cigar;
    std::vector<ShimRead> allreads(1);
    allreads[0].neighbors_.push_back(0);
    const int backbone = 0;
    std::pair<std::pair<int, int>, std::vector<int>> interval;
    interval.second.push_back(0);
    for (int r = 0; r < 1; r++) {
#include "_ref/clip_snippet.inc"
        out[0] = posOnReadStart;
        out[1] = posOnReadEnd;
        out[2] = posOnCIGARStart;
        out[3] = posOnCIGAREnd;
        (void)startPosition;
        return 0;
    }
    return interval.second[0];
}

// convert_cigar2(converted_cigar.substr(a, b - a)) (:459-461): the clipped CIGAR string
int hsref_clip_cigar(const char* cigar, int a, int b, char* out, int capacity) {
    std::string c = cigar;
    std::string conv = convert_cigar(c);
    std::string sub = conv.substr(a, b - a);
    std::string res = convert_cigar2(sub);
    if ((int)res.size() + 1 > capacity) return -1;
    std::copy(res.begin(), res.end(), out);
    out[res.size()] = 0;
    return (int)res.size();
}
}
