// moira_collapse.cpp -- host-side dereplication of identical sequences, the step that
// sits right behind the filter kernels when --collapse is on.
//
// Reference semantics (moira/moira.py:459-475 and the epilogue :491-504), reproduced exactly:
//   * reads with the same (already truncated) sequence string form one group, groups are numbered in
//     order of first appearance (a Python-3 dict's insertion order);
//   * walking a group's reads in input order, the representative starts as the first read and is
//     replaced by a later read only if that read's expected errors are STRICTLY smaller (:466); every
//     new representative is inserted at the FRONT of the names list (:470), every other read is appended;
//   * groups are written by abundance, largest first (`sorted(..., reverse=True)`, :492 -- stable, so
//     ties keep first-appearance order).
//
// Parallel plan: 64-bit hash of every sequence (all host threads) -> stable counting sort of the read
// indices into 256 hash buckets -> each bucket grouped independently with an open-addressing table
// (hash match + memcmp: exact strings, never hashes alone) -> groups ordered by first read.
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "moira_internal.h"

namespace {

inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    x *= 0xd6e8feb86659fd93ull;
    x ^= x >> 32;
    return x;
}

uint64_t hash_bytes(const char *p, uint32_t len)
{
    uint64_t h = 0x9e3779b97f4a7c15ull ^ len;
    uint32_t i = 0;
    for (; i + 8 <= len; i += 8) {
        uint64_t w;
        memcpy(&w, p + i, 8);
        h = mix64(h ^ w) + 0x9e3779b97f4a7c15ull;
    }
    uint64_t w = 0;
    if (i < len) memcpy(&w, p + i, len - i);
    return mix64(h ^ w);
}

struct LocalGroup {
    uint64_t first;                 // first read of the group (also its hash-table identity)
    std::vector<uint64_t> members;  // in input order
};

template <typename F>
void parallel_for(int n_items, int threads, F &&fn)
{
    moira::parallel_run(n_items, threads, std::function<void(int)>(fn));
}

}  // namespace

// abundance_order = the groups by size, largest first, equal sizes in order of first appearance (sorted(..., reverse=True)
// on the sizes is stable, moira.py:492): a counting sort -- sizes are small integers, a comparison sort of millions of
// groups cost more than everything else in the collapse.
static void sort_by_abundance(const uint64_t *group_size, uint64_t G, uint64_t *abundance_order)
{
    uint64_t max_size = 0;
    for (uint64_t g = 0; g < G; g++) max_size = std::max(max_size, group_size[g]);
    if (max_size > (1ull << 26)) {   // absurdly large groups: fall back to the comparison sort
        for (uint64_t g = 0; g < G; g++) abundance_order[g] = g;
        std::stable_sort(abundance_order, abundance_order + G, [&](uint64_t x, uint64_t y) { return group_size[x] > group_size[y]; });
        return;
    }
    std::vector<uint64_t> start(max_size + 2, 0);
    for (uint64_t g = 0; g < G; g++) start[group_size[g]]++;
    uint64_t pos = 0;
    for (uint64_t sz = max_size + 1; sz-- > 0;) { const uint64_t c = start[sz]; start[sz] = pos; pos += c; }
    for (uint64_t g = 0; g < G; g++) abundance_order[start[group_size[g]]++] = g;
}

extern "C" int moira_collapse(const char *text, const uint64_t *seq_off, const uint32_t *seq_len, const double *ee,
                              uint64_t n, int n_threads, uint64_t *group_of_read, uint64_t *n_groups_out,
                              uint64_t *group_rep, uint64_t *group_size, uint64_t *member_start, uint64_t *members,
                              uint64_t *abundance_order)
{
    if ((n && (!seq_off || !seq_len || !ee || !group_of_read || !group_rep || !group_size || !member_start || !members ||
               !abundance_order)) || !n_groups_out)
        return moira::fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    // text == NULL: seq_off holds absolute addresses (sequences spread over several buffers)
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > 64) T = 64;
    constexpr int NBK = 256;
    std::vector<uint64_t> hash(n);
    {
        const int parts = T * 4;
        parallel_for(parts, T, [&](int p) {
            for (uint64_t r = n * (uint64_t)p / parts, e = n * (uint64_t)(p + 1) / parts; r < e; r++)
                hash[r] = hash_bytes((const char *)((uintptr_t)text + seq_off[r]), seq_len[r]);
        });
    }
    // stable counting sort of read indices by the top 8 hash bits
    std::vector<uint64_t> bstart(NBK + 1, 0), sorted(n);
    for (uint64_t r = 0; r < n; r++) bstart[(hash[r] >> 56) + 1]++;
    for (int b = 0; b < NBK; b++) bstart[b + 1] += bstart[b];
    {
        std::vector<uint64_t> cur(bstart.begin(), bstart.end() - 1);
        for (uint64_t r = 0; r < n; r++) sorted[cur[hash[r] >> 56]++] = r;
    }
    // group every bucket on its own
    std::vector<std::vector<LocalGroup>> local(NBK);
    parallel_for(NBK, T, [&](int b) {
        const uint64_t lo = bstart[b], hi = bstart[b + 1], cnt = hi - lo;
        if (!cnt) return;
        uint64_t cap = 16;
        while (cap < cnt * 2) cap <<= 1;
        std::vector<uint32_t> table(cap, 0xFFFFFFFFu);     // slot -> local group index
        std::vector<LocalGroup> &groups = local[b];
        for (uint64_t k = lo; k < hi; k++) {
            const uint64_t r = sorted[k];
            const uint64_t h = hash[r];
            uint64_t slot = (h >> 8) & (cap - 1);
            for (;;) {
                const uint32_t gi = table[slot];
                if (gi == 0xFFFFFFFFu) {
                    table[slot] = (uint32_t)groups.size();
                    groups.push_back(LocalGroup{r, {r}});
                    break;
                }
                const uint64_t f = groups[gi].first;
                if (hash[f] == h && seq_len[f] == seq_len[r] && memcmp((const char *)((uintptr_t)text + seq_off[f]), (const char *)((uintptr_t)text + seq_off[r]), seq_len[r]) == 0) {
                    groups[gi].members.push_back(r);
                    break;
                }
                slot = (slot + 1) & (cap - 1);
            }
        }
    });
    // global group ids in order of first appearance
    struct Ref { uint64_t first; int bucket; uint32_t idx; };
    std::vector<Ref> refs;
    for (int b = 0; b < NBK; b++)
        for (uint32_t i = 0; i < local[b].size(); i++) refs.push_back(Ref{local[b][i].first, b, i});
    std::sort(refs.begin(), refs.end(), [](const Ref &x, const Ref &y) { return x.first < y.first; });
    const uint64_t G = refs.size();
    *n_groups_out = G;
    uint64_t pos = 0;
    for (uint64_t g = 0; g < G; g++) {
        member_start[g] = pos;
        pos += local[refs[g].bucket][refs[g].idx].members.size();
    }
    if (n) member_start[G] = pos;
    // representative + names order of every group (moira.py:466-475)
    {
        const int parts = T * 8;
        parallel_for(parts, T, [&](int p) {
            std::vector<uint64_t> breakers, others;
            for (uint64_t g = G * (uint64_t)p / parts, e = G * (uint64_t)(p + 1) / parts; g < e; g++) {
                const std::vector<uint64_t> &m = local[refs[g].bucket][refs[g].idx].members;
                breakers.clear();
                others.clear();
                uint64_t rep = m[0];
                breakers.push_back(rep);
                for (size_t k = 1; k < m.size(); k++) {
                    if (ee[m[k]] < ee[rep]) { rep = m[k]; breakers.push_back(rep); }
                    else others.push_back(m[k]);
                }
                uint64_t *out = members + member_start[g];
                for (size_t k = 0; k < breakers.size(); k++) out[k] = breakers[breakers.size() - 1 - k];
                for (size_t k = 0; k < others.size(); k++) out[breakers.size() + k] = others[k];
                group_rep[g] = rep;
                group_size[g] = m.size();
                for (uint64_t r : m) group_of_read[r] = g;
            }
        });
    }
    sort_by_abundance(group_size, G, abundance_order);
    return MOIRA_OK;
}

// The same outputs from labels (equal label <=> equal sequence, label < n) -- the labels come from the device-side
// dereplication (moira_dedup.cu), so no base is looked at here: three linear passes over the reads in input order,
// which is exactly the order the reference's dictionary sees them in (moira.py:455-475).
extern "C" int moira_collapse_labels(const uint32_t *labels, const double *ee, uint64_t n, uint64_t *group_of_read, uint64_t *n_groups_out,
                                     uint64_t *group_rep, uint64_t *group_size, uint64_t *member_start, uint64_t *members,
                                     uint64_t *abundance_order)
{
    if ((n && (!labels || !ee || !group_of_read || !group_rep || !group_size || !member_start || !members || !abundance_order)) ||
        !n_groups_out)
        return moira::fail(MOIRA_ERR_BAD_ARG, "NULL argument");
    constexpr uint64_t NONE = ~0ull;
    std::vector<uint64_t> gid(n, NONE);       // by label
    uint64_t G = 0;
    for (uint64_t r = 0; r < n; r++) {        // groups in order of first appearance
        const uint32_t l = labels[r];
        if (l >= n) return moira::fail(MOIRA_ERR_BAD_ARG, "label %u of read %llu is not a read index", l, (unsigned long long)r);
        uint64_t g = gid[l];
        if (g == NONE) { g = G++; gid[l] = g; group_size[g] = 0; group_rep[g] = r; }
        group_of_read[r] = g;
        group_size[g]++;
    }
    *n_groups_out = G;
    // representative = first read with the strictly smallest ee (:466); count the record breakers on the way
    std::vector<uint32_t> breakers(G, 0);
    {
        std::vector<uint8_t> seen(G, 0);
        for (uint64_t r = 0; r < n; r++) {
            const uint64_t g = group_of_read[r];
            if (!seen[g]) { seen[g] = 1; group_rep[g] = r; breakers[g] = 1; }
            else if (ee[r] < ee[group_rep[g]]) { group_rep[g] = r; breakers[g]++; }
        }
    }
    uint64_t pos = 0;
    for (uint64_t g = 0; g < G; g++) { member_start[g] = pos; pos += group_size[g]; }
    if (n) member_start[G] = pos;
    // names order (:470-475): every new representative goes to the FRONT, every other read to the back
    {
        std::vector<uint64_t> front(G), back(G), cur(G, NONE);
        for (uint64_t g = 0; g < G; g++) { front[g] = member_start[g] + breakers[g]; back[g] = front[g]; }
        for (uint64_t r = 0; r < n; r++) {
            const uint64_t g = group_of_read[r];
            if (cur[g] == NONE || ee[r] < ee[cur[g]]) { cur[g] = r; members[--front[g]] = r; }
            else members[back[g]++] = r;
        }
    }
    sort_by_abundance(group_size, G, abundance_order);
    return MOIRA_OK;
}
