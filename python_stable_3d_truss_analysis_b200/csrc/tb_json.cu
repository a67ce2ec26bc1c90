// Bulk loader of the reference's truss JSON files (host code only).
//
// Truss.LoadFromJSON (slientruss3d/truss.py:401-421) reads one file into one Truss object through AddNewJoint /
// AddExternalForce / AddNewMember, a Python call per joint, load and member.  A dataset is thousands of such files
// (generate.py:361-362 writes one per truss), and what the solver wants is the packed arrays of tb_ragged_in.  This file
// parses N documents straight into those arrays on a pool of host threads: a schema-driven recursive-descent scanner over
// the text (no DOM), numbers through strtod (correctly rounded, the same double json.load produces), the reference's
// semantics for repeated / zero load vectors (truss.py:177-182) and for the sparse result lists of an output file
// (truss.py:414-418: entries the reference dropped under its 1e-10 filter come back as zeros).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <cmath>
#include <thread>
#include <vector>

#include "truss_b200.h"

namespace {

struct Cur {
  const char* p;
  const char* end;
  bool ok = true;
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) { ++p; return true; }
    return false;
  }
  bool expect(char c) {
    if (eat(c)) return true;
    ok = false;
    return false;
  }
  char peek() {
    ws();
    return p < end ? *p : '\0';
  }
  // string -> [s, e) without the quotes (escapes are skipped over, not decoded: keys and support names have none)
  bool str(const char** s, const char** e) {
    ws();
    if (p >= end || *p != '"') { ok = false; return false; }
    ++p;
    *s = p;
    while (p < end && *p != '"') {
      if (*p == '\\') ++p;
      ++p;
    }
    if (p >= end) { ok = false; return false; }
    *e = p;
    ++p;
    return true;
  }
  bool num(double* v) {
    ws();
    if (p >= end) { ok = false; return false; }
    // (the documents are NUL-terminated by the caller, so strtod cannot run past the buffer)
    char* q = nullptr;
    const double x = strtod(p, &q);
    if (q == p) {
      // json.load also accepts the bare tokens NaN / Infinity / -Infinity; strtod wants them spelled its own way
      if (end - p >= 3 && !strncmp(p, "NaN", 3)) { *v = NAN; p += 3; return true; }
      ok = false;
      return false;
    }
    p = q;
    *v = x;
    return true;
  }
  void skip_value() {  // any JSON value
    ws();
    if (p >= end) { ok = false; return; }
    const char c = *p;
    if (c == '"') {
      const char *s, *e;
      str(&s, &e);
    } else if (c == '{' || c == '[') {
      const char close = c == '{' ? '}' : ']';
      ++p;
      if (eat(close)) return;
      do {
        if (c == '{') {
          const char *s, *e;
          if (!str(&s, &e) || !expect(':')) return;
        }
        skip_value();
        if (!ok) return;
      } while (eat(','));
      expect(close);
    } else if (c == 't' || c == 'f' || c == 'n') {
      while (p < end && *p >= 'a' && *p <= 'z') ++p;
    } else {
      double v;
      num(&v);
    }
  }
};

bool key_is(const char* s, const char* e, const char* lit) {
  const size_t n = strlen(lit);
  return (size_t)(e - s) == n && !memcmp(s, lit, n);
}

int support_code(const char* s, const char* e) {   // type.py:5-10 names -> the ABI codes (TB_SUPPORT_*)
  if (key_is(s, e, "NO")) return 0;
  if (key_is(s, e, "PIN")) return 1;
  if (key_is(s, e, "ROLLER_X")) return 2;
  if (key_is(s, e, "ROLLER_Y")) return 3;
  if (key_is(s, e, "ROLLER_Z")) return 4;
  return -1;
}

// [v0, v1, ...]: the first `dim` components are kept (AddNewJoint / AddExternalForce read vector[i], i < dim)
bool vec(Cur& c, int dim, double* out) {
  if (!c.expect('[')) return false;
  int i = 0;
  if (c.peek() != ']') {
    do {
      double v;
      if (!c.num(&v)) return false;
      if (i < dim) out[i] = v;
      ++i;
    } while (c.eat(','));
  }
  if (!c.expect(']')) return false;
  if (i < dim) { c.ok = false; return false; }
  return true;
}

struct Doc {
  int64_t n_joint = 0, n_member = 0;
};

// pass 1: count joints and members
int scan_doc(const char* text, int64_t len, Doc* d) {
  Cur c{text, text + len};
  if (!c.expect('{')) return TB_ERR_JSON;
  if (c.peek() != '}') {
    do {
      const char *ks, *ke;
      if (!c.str(&ks, &ke) || !c.expect(':')) return TB_ERR_JSON;
      const bool isj = key_is(ks, ke, "joint"), ism = key_is(ks, ke, "member");
      if (isj || ism) {
        int64_t n = 0;
        if (!c.expect('[')) return TB_ERR_JSON;
        if (c.peek() != ']') {
          do {
            c.skip_value();
            if (!c.ok) return TB_ERR_JSON;
            ++n;
          } while (c.eat(','));
        }
        if (!c.expect(']')) return TB_ERR_JSON;
        (isj ? d->n_joint : d->n_member) = n;
      } else {
        c.skip_value();
        if (!c.ok) return TB_ERR_JSON;
      }
    } while (c.eat(','));
  }
  if (!c.expect('}')) return TB_ERR_JSON;
  return TB_OK;
}

struct Fill {
  int dim, is_output;
  int64_t nJ, nM;
  double* xyz; uint8_t* support; double* force; int32_t* conn; double* aed;
  double* u; double* ext; double* axial; double* weight;
};

// [[id, [vector]], ...] -> rows of a dense [nJ][dim] array; `skip_zero`: AddExternalForce ignores zero vectors (truss.py:181)
int joint_list(Cur& c, const Fill& f, double* dst, bool skip_zero) {
  if (!c.expect('[')) return TB_ERR_JSON;
  if (c.peek() != ']') {
    do {
      double id, v[3] = {0, 0, 0};
      if (!c.expect('[') || !c.num(&id) || !c.expect(',') || !vec(c, f.dim, v) || !c.expect(']')) return TB_ERR_JSON;
      const int64_t j = (int64_t)id;
      if ((double)j != id || j < 0 || j >= f.nJ) return TB_ERR_INDEX;
      bool zero = true;
      for (int i = 0; i < f.dim; ++i) zero = zero && std::fabs(v[i]) < 1e-10;
      if (dst && !(skip_zero && zero))
        for (int i = 0; i < f.dim; ++i) dst[j * f.dim + i] = v[i];
    } while (c.eat(','));
  }
  return c.expect(']') ? TB_OK : TB_ERR_JSON;
}

// pass 2: fill this document's slices
int fill_doc(const char* text, int64_t len, const Fill& f) {
  Cur c{text, text + len};
  const int d = f.dim;
  bool have_j = false, have_m = false, have_f = false;
  bool res[3] = {false, false, false};
  if (!c.expect('{')) return TB_ERR_JSON;
  if (c.peek() != '}') {
    do {
      const char *ks, *ke;
      if (!c.str(&ks, &ke) || !c.expect(':')) return TB_ERR_JSON;
      if (key_is(ks, ke, "joint")) {                 // [[x, y, z], "SUPPORT"]   (truss.py:406-407)
        int64_t j = 0;
        if (!c.expect('[')) return TB_ERR_JSON;
        if (c.peek() != ']') {
          do {
            if (j >= f.nJ) return TB_ERR_JSON;
            const char *ss, *se;
            if (!c.expect('[') || !vec(c, d, f.xyz + j * d) || !c.expect(',') || !c.str(&ss, &se) || !c.expect(']'))
              return TB_ERR_JSON;
            const int code = support_code(ss, se);
            if (code < 0) return TB_ERR_SUPPORT;
            f.support[j] = (uint8_t)code;
            ++j;
          } while (c.eat(','));
        }
        if (!c.expect(']') || j != f.nJ) return TB_ERR_JSON;
        have_j = true;
      } else if (key_is(ks, ke, "force")) {          // [jointID, [fx, fy, fz]]  (truss.py:409-410)
        const int rc = joint_list(c, f, f.force, true);     // (ids are checked against the joint count of pass 1)
        if (rc) return rc;
        have_f = true;
      } else if (key_is(ks, ke, "member")) {         // [[j0, j1], [a, e, density]]  (truss.py:412-413)
        int64_t m = 0;
        if (!c.expect('[')) return TB_ERR_JSON;
        if (c.peek() != ']') {
          do {
            if (m >= f.nM) return TB_ERR_JSON;
            double jj[3] = {0, 0, 0};
            if (!c.expect('[')) return TB_ERR_JSON;
            if (!c.expect('[') || !c.num(&jj[0]) || !c.expect(',') || !c.num(&jj[1]) || !c.expect(']') || !c.expect(','))
              return TB_ERR_JSON;
            // member type: exactly three numbers (MemberType(*memberType))
            if (!c.expect('[') || !c.num(&f.aed[3 * m]) || !c.expect(',') || !c.num(&f.aed[3 * m + 1]) || !c.expect(',') ||
                !c.num(&f.aed[3 * m + 2]) || !c.expect(']') || !c.expect(']'))
              return TB_ERR_JSON;
            for (int e = 0; e < 2; ++e) {
              const int64_t j = (int64_t)jj[e];
              if ((double)j != jj[e] || j < 0 || j >= f.nJ) return TB_ERR_INDEX;
              f.conn[2 * m + e] = (int32_t)j;
            }
            ++m;
          } while (c.eat(','));
        }
        if (!c.expect(']') || m != f.nM) return TB_ERR_JSON;
        have_m = true;
      } else if (f.is_output && key_is(ks, ke, "displace")) {
        const int rc = joint_list(c, f, f.u, false);
        if (rc) return rc;
        res[0] = true;
      } else if (f.is_output && key_is(ks, ke, "external")) {
        const int rc = joint_list(c, f, f.ext, false);
        if (rc) return rc;
        res[1] = true;
      } else if (f.is_output && key_is(ks, ke, "internal")) {   // [memberID, force]
        if (!c.expect('[')) return TB_ERR_JSON;
        if (c.peek() != ']') {
          do {
            double id, v;
            if (!c.expect('[') || !c.num(&id) || !c.expect(',') || !c.num(&v) || !c.expect(']')) return TB_ERR_JSON;
            const int64_t m = (int64_t)id;
            if ((double)m != id || m < 0 || m >= f.nM) return TB_ERR_INDEX;
            if (f.axial) f.axial[m] = v;
          } while (c.eat(','));
        }
        if (!c.expect(']')) return TB_ERR_JSON;
        res[2] = true;
      } else if (f.is_output && key_is(ks, ke, "weight")) {
        double w;
        if (!c.num(&w)) return TB_ERR_JSON;
        if (f.weight) *f.weight = w;
      } else {
        c.skip_value();
        if (!c.ok) return TB_ERR_JSON;
      }
    } while (c.eat(','));
  }
  if (!c.expect('}')) return TB_ERR_JSON;
  (void)have_f;
  if (!have_j || !have_m) return TB_ERR_JSON;       // LoadFromJSON raises KeyError without them
  if (f.is_output && !(res[0] && res[1] && res[2])) return TB_ERR_JSON;
  return TB_OK;
}

template <class F>
void parallel_for(int n, int threads, F&& body) {
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  if (threads <= 1) {
    for (int i = 0; i < n; ++i) body(i);
    return;
  }
  std::atomic<int> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] {
      for (;;) {
        const int i0 = next.fetch_add(16);
        if (i0 >= n) break;
        for (int i = i0; i < n && i < i0 + 16; ++i) body(i);
      }
    });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int tb_json_scan(int32_t n, const char* const* texts, const int64_t* lens, int64_t* n_joint, int64_t* n_member,
                            int32_t* err, int32_t threads) {
  if (n < 0) return TB_ERR_SIZE;
  if (n == 0) return TB_OK;
  if (!texts || !lens || !n_joint || !n_member) return TB_ERR_NULL;
  std::atomic<int> bad{0};
  parallel_for(n, threads, [&](int i) {
    Doc d;
    const int rc = texts[i] ? scan_doc(texts[i], lens[i], &d) : TB_ERR_NULL;
    n_joint[i] = d.n_joint;
    n_member[i] = d.n_member;
    if (err) err[i] = rc;
    if (rc) bad.store(1);
  });
  return bad.load() ? TB_ERR_JSON : TB_OK;
}

extern "C" int tb_json_fill(int32_t n, const char* const* texts, const int64_t* lens, int32_t dim, int32_t is_output,
                            const int64_t* joint_off, const int64_t* member_off, double* xyz, uint8_t* support, double* force,
                            int32_t* conn, double* aed, double* u, double* ext, double* axial, double* weight, int32_t* err,
                            int32_t threads) {
  if (dim != 2 && dim != 3) return TB_ERR_DIM;
  if (n < 0) return TB_ERR_SIZE;
  if (n == 0) return TB_OK;
  if (!texts || !lens || !joint_off || !member_off || !xyz || !support || !force || !conn || !aed) return TB_ERR_NULL;
  if (is_output && (!u || !ext || !axial)) return TB_ERR_NULL;
  std::atomic<int> bad{0};
  parallel_for(n, threads, [&](int i) {
    Fill f;
    f.dim = dim;
    f.is_output = is_output;
    const int64_t j0 = joint_off[i], m0 = member_off[i];
    f.nJ = joint_off[i + 1] - j0;
    f.nM = member_off[i + 1] - m0;
    f.xyz = xyz + j0 * dim;
    f.support = support + j0;
    f.force = force + j0 * dim;
    f.conn = conn + 2 * m0;
    f.aed = aed + 3 * m0;
    f.u = is_output ? u + j0 * dim : nullptr;
    f.ext = is_output ? ext + j0 * dim : nullptr;
    f.axial = is_output ? axial + m0 : nullptr;
    f.weight = is_output && weight ? weight + i : nullptr;
    // loads and sparse results default to zero (a joint without a "force" entry carries no load)
    memset(f.force, 0, sizeof(double) * f.nJ * dim);
    if (is_output) {
      memset(f.u, 0, sizeof(double) * f.nJ * dim);
      memset(f.ext, 0, sizeof(double) * f.nJ * dim);
      memset(f.axial, 0, sizeof(double) * f.nM);
      if (f.weight) *f.weight = NAN;
    }
    const int rc = texts[i] ? fill_doc(texts[i], lens[i], f) : TB_ERR_NULL;
    if (err) err[i] = rc;
    if (rc) bad.store(1);
  });
  return bad.load() ? TB_ERR_JSON : TB_OK;
}
