// Host-only C ABI over the symbolic analysis (no CUDA calls): lets the CPU test
// suite check ordering, fill, level sets and schedules without a GPU.
#include <cstring>
#include <string>

#include "../../include/helm_b200.h"
#include "symbolic.h"

struct helm_symbolic {
    helm::Symbolic S;
    std::string err;
};

extern "C" {

int helm_symbolic_create(int32_t n_bus, const int32_t *adj_ptr, const int32_t *adj_idx,
                         const uint8_t *bus_type, helm_symbolic **out) {
    if (!adj_ptr || !adj_idx || !out) return HELM_E_ARG;
    helm_symbolic *h = new helm_symbolic();
    if (!helm::build_symbolic(n_bus, adj_ptr, adj_idx, bus_type, h->S, h->err)) {
        delete h;
        return HELM_E_STRUCT;
    }
    *out = h;
    return HELM_OK;
}

void helm_symbolic_destroy(helm_symbolic *h) { delete h; }

int helm_symbolic_array(const helm_symbolic *h, const char *name, const int32_t **data, int32_t *len) {
    if (!h || !name || !data || !len) return HELM_E_ARG;
    const helm::Symbolic &S = h->S;
    const std::vector<int> *v = nullptr;
#define F(n) if (!strcmp(name, #n)) v = &S.n;
    F(perm) F(iperm) F(adj_ptr) F(adj_idx) F(adj_src) F(adj_slot) F(lptr) F(lcol) F(llev_ptr)
    F(urow) F(qof) F(ulev_ptr) F(uptr) F(ucol) F(dslot) F(fptr) F(fsrc) F(fdst) F(fdst_local) F(fchunk_row) F(fchunk_lpr) F(flev_chunk) F(frow_off) F(slot_src) F(slot_row) F(slot_col) F(wave_tab) F(smeta) F(scm) F(solve_chunks) F(solve_segs) F(one_lists) F(rl_desc) F(rc_lists) F(rc_segend) F(rc_desc) F(rc_seg)
#undef F
    if (!v) return HELM_E_ARG;
    *data = v->data();
    *len = (int32_t)v->size();
    return HELM_OK;
}

}  // extern "C"
