/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Calls the reference's OWN UMI-grouping entry points -- umi_group (src/umi_group.cpp:14-117), fast_levdist_test
 * (src/sorted_trie.cpp:302-332) and cluster_umis_test (src/cluster_umis_test.cpp:8-29) -- compiled verbatim from
 * /root/reference/src against the toy R object model of oracle/rshim/.  Nothing is restated here: this file only
 * builds the argument objects and flattens the returned list of integer vectors.
 * Results are kept in a static holder and fetched with a second call (sizes are not known up front).  Not thread-safe.
 */
#include "sarlacc.h"

#include <cstring>
#include <cstdint>

RshimObject rshim_nil_object;
std::vector<std::unique_ptr<RshimObject> >& rshim_arena() {
    static std::vector<std::unique_ptr<RshimObject> > arena;
    return arena;
}
std::string& rshim_last_error() {
    static std::string msg;
    return msg;
}

extern "C" {
/* Biostrings' interface: the S4 path is never taken by the oracle */
XStringSet_holder hold_XStringSet(struct RshimObject*) { throw std::runtime_error("rshim: no XStringSet support"); }
int get_length_from_XStringSet_holder(const XStringSet_holder* x) { return x->length; }
Chars_holder get_elt_from_XStringSet_holder(const XStringSet_holder*, int) { throw std::runtime_error("rshim: no XStringSet support"); }
char DNAdecode(char code) { return code; }
}

namespace {

std::vector<std::vector<int> > g_result;

SEXP make_strings(const char* pool, const int64_t* off, int64_t n) {
    SEXP s = rshim_new(RSHIM_STR);
    s->strs.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) s->strs[(size_t)i].assign(pool + off[i], pool + off[i + 1]);
    return s;
}

SEXP make_int(int v) {
    SEXP s = rshim_new(RSHIM_INT);
    s->ints.assign(1, v);
    return s;
}

SEXP make_int_list(const int64_t* off, const int32_t* values, int64_t n) {
    SEXP l = rshim_new(RSHIM_LIST);
    for (int64_t g = 0; g < n; ++g) {
        SEXP v = rshim_new(RSHIM_INT);
        v->ints.assign(values + off[g], values + off[g + 1]);
        l->items.push_back(v);
    }
    return l;
}

/* list of integer vectors, or (umi_group) list of lists of integer vectors, flattened one level like
 * unlist(out, recursive=FALSE) in R/umiGroup.R:22 */
void flatten(SEXP out, bool nested) {
    g_result.clear();
    for (SEXP a : out->items) {
        if (nested) {
            for (SEXP b : a->items) g_result.push_back(b->ints);
        } else {
            g_result.push_back(a->ints);
        }
    }
}

int finish(SEXP out, bool nested, int64_t* n_lists, int64_t* n_values, char* err, int errlen) {
    int rc = 0;
    if (!out) {
        if (err && errlen > 0) {
            std::strncpy(err, rshim_last_error().c_str(), errlen - 1);
            err[errlen - 1] = '\0';
        }
        g_result.clear();
        rc = 1;
    } else {
        flatten(out, nested);
    }
    int64_t tot = 0;
    for (auto& v : g_result) tot += (int64_t)v.size();
    if (n_lists) *n_lists = (int64_t)g_result.size();
    if (n_values) *n_values = tot;
    rshim_arena().clear();
    return rc;
}

}

extern "C" {

/* .Call(cxx_umi_group, UMI1, threshold1, UMI2, threshold2, by.group) + unlist(recursive=FALSE), R/umiGroup.R:21-22.
 * members are 1-based, as R passes them. */
int ref_umi_group(const char* pool1, const int64_t* off1, int64_t n, int thresh1,
                  const char* pool2, const int64_t* off2, int thresh2,
                  const int64_t* group_off, const int32_t* members, int64_t ngroups,
                  int64_t* n_lists, int64_t* n_values, char* err, int errlen)
{
    SEXP u1 = make_strings(pool1, off1, n);
    SEXP u2 = pool2 ? make_strings(pool2, off2, n) : R_NilValue;
    SEXP out = umi_group(u1, make_int(thresh1), u2, make_int(thresh2), make_int_list(group_off, members, ngroups));
    return finish(out, true, n_lists, n_values, err, errlen);
}

/* .Call(cxx_fast_levdist_test, seqs, limit, sorted): 1-based neighbour lists in trie order */
int ref_levdist(const char* pool, const int64_t* off, int64_t n, int limit, int sorted,
                int64_t* n_lists, int64_t* n_values, char* err, int errlen)
{
    SEXP lg = rshim_new(RSHIM_LGL);
    lg->ints.assign(1, sorted ? 1 : 0);
    SEXP out = fast_levdist_test(make_strings(pool, off, n), make_int(limit), lg);
    return finish(out, false, n_lists, n_values, err, errlen);
}

/* .Call(cxx_cluster_umis_test, links): links and result 1-based */
int ref_cluster_umis(const int64_t* off, const int32_t* links, int64_t n,
                     int64_t* n_lists, int64_t* n_values, char* err, int errlen)
{
    SEXP out = cluster_umis_test(make_int_list(off, links, n));
    return finish(out, false, n_lists, n_values, err, errlen);
}

void ref_fetch(int64_t* list_off, int32_t* values) {
    int64_t at = 0;
    for (size_t i = 0; i < g_result.size(); ++i) {
        list_off[i] = at;
        for (int v : g_result[i]) values[at++] = v;
    }
    list_off[g_result.size()] = at;
}

}
