/* oracle/krepp_oracle.c -- TEST INFRASTRUCTURE ONLY (see krepp_oracle.h).
 *
 * Plain-C restatement of the krepp v0.8.3 query path.  Nothing here is used by the product (krepp_b200/); it is the
 * checker.  All "ref:" citations are relative to /root/reference.  Arithmetic is kept in the reference's literal
 * evaluation order (no FMA contraction: built without -march/-ffast-math, like the reference).
 */
#define _GNU_SOURCE
#include "krepp_oracle.h"
#include <ctype.h>
#include <dirent.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

/* ------------------------------------------------------------------------------------------------ data model */

typedef struct {
  /* one partial table: ref src/table.hpp:139-145 (FlatHT), src/record.hpp:97-104 (CRecord) */
  uint32_t r, frac, nrows;
  uint64_t nkmers;
  uint32_t* cmer; /* nkmers x {enc, se} */
  uint64_t* inc;  /* nrows cumulative ends */
  uint32_t nnodes, nsubsets;
  uint32_t* pse; /* nsubsets x {first, second} */
  double* rho;   /* nnodes */
} ko_table_t;

typedef struct {
  uint32_t nnodes;      /* tree nodes; se in 1..nnodes, 0 = null sentinel (ref src/phytree.hpp:53) */
  uint32_t* parent;     /* [nnodes+1], 0 for root */
  uint32_t* nchildren;  /* [nnodes+1] */
  uint32_t* eff_nch;    /* [nnodes+1]  == nchildren unless a query tree was mapped (not restated) */
  uint8_t* is_leaf;     /* [nnodes+1] */
  uint32_t* card;       /* [nnodes+1] leaves below */
  double* blen;         /* [nnodes+1] NaN when absent */
  char** name;          /* [nnodes+1] "" when unlabeled */
  uint32_t** children;  /* [nnodes+1] child se lists, in Newick order */
  uint32_t root;
} ko_tree_t;

struct ko_index {
  uint32_t k, w, h, m;
  uint8_t ppos[32], npos[32];
  uint64_t mask_hash_bp, mask_drop_lr, mask_bp, mask_lr;
  uint32_t geom_r, geom_frac;  /* geometry-only handles (ko_geom_new) */
  ko_table_t* tables; uint32_t ntables;
  int32_t* res_table;          /* [m] residue -> table index or -1   (ref src/index.cpp:144-157 r_to_flatht) */
  uint32_t* res_numer;         /* [m] residue -> numerator           (ref r_to_numerator) */
  ko_tree_t tree;
  int have_tree;
};

static void seterr(char* err, size_t n, const char* fmt, ...)
{
  if (!err || !n) return;
  va_list ap; va_start(ap, fmt); vsnprintf(err, n, fmt, ap); va_end(ap);
}

/* ------------------------------------------------------------------------------------------------ primitives */

/* ref src/common.cpp:10-14 (seq_nt4_table): A/a 0, C/c 1, G/g 2, T/t 3, everything else 4.  The reference indexes a
 * 128-entry table with a plain char; bytes >= 128 are out of its domain and are treated as non-ACGT here. */
static inline unsigned nt4(unsigned char c)
{
  switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
               case 'T': case 't': return 3; default: return 4; }
}

/* ref src/lshf.cpp:62-71: the hash is defined as _pext_u64 (the shld and extract_bits variants are equivalents). */
uint64_t ko_pext64(uint64_t x, uint64_t mask)
{
  uint64_t res = 0, bb = 1;
  for (; mask; mask &= mask - 1, bb += bb)
    if (x & mask & (~mask + 1)) res |= bb;
  return res;
}

/* ref src/common.hpp:177-186 */
uint64_t ko_revcomp_bp64(uint64_t x, uint32_t k)
{
  uint64_t res = ~x;
  res = ((res >> 2 & 0x3333333333333333ull) | (res & 0x3333333333333333ull) << 2);
  res = ((res >> 4 & 0x0F0F0F0F0F0F0F0Full) | (res & 0x0F0F0F0F0F0F0F0Full) << 4);
  res = ((res >> 8 & 0x00FF00FF00FF00FFull) | (res & 0x00FF00FF00FF00FFull) << 8);
  res = ((res >> 16 & 0x0000FFFF0000FFFFull) | (res & 0x0000FFFF0000FFFFull) << 16);
  res = ((res >> 32 & 0x00000000FFFFFFFFull) | (res & 0x00000000FFFFFFFFull) << 32);
  return res >> (2 * (32 - k));
}

/* ref src/common.hpp:188-197 (rmoddp_bp64) and :223 (conv_bp64_lr64) */
static uint64_t rmoddp(uint64_t x)
{
  x = x & 0x5555555555555555ull;
  x = (x | (x >> 1)) & 0x3333333333333333ull;
  x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
  x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
  x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
  x = (x | (x >> 16)) & 0x00000000ffffffffull;
  return x;
}
uint64_t ko_conv_bp64_lr64(uint64_t x) { return (rmoddp(x >> 1) << 32) | rmoddp(x); }

/* ref src/common.hpp:175 */
uint32_t ko_popcount_lr32(uint32_t z) { return (uint32_t)__builtin_popcount((z | (z >> 16)) & 0x0000ffffu); }

/* ref src/common.hpp:147-155 */
uint64_t ko_xur64_hash(uint64_t h)
{
  h ^= (h >> 33); h *= 0xff51afd7ed558ccdull; h ^= (h >> 33); h *= 0xc4ceb9fe1a85ec53ull; h ^= (h >> 33);
  return h;
}

/* ref src/lshf.cpp:39-54 (set_lshf masks) and src/query.cpp:35-37 (k-mer masks) */
static void set_masks(ko_index_t* ix)
{
  uint32_t k = ix->k, h = ix->h;
  ix->mask_drop_lr = 0; ix->mask_hash_bp = 0;
  for (int i = (int)(k - h) - 1; i >= 0; --i) ix->mask_drop_lr += (0x0000000100000001ull << ix->npos[i]);
  for (uint32_t i = 0; i < 16 - (k - h); ++i) ix->mask_drop_lr += 0x0000000000000001ull << (i + k);
  for (int i = (int)h - 1; i >= 0; --i) ix->mask_hash_bp += (0x0000000000000003ull << (ix->ppos[i] * 2));
  uint64_t u64m = ~0ull;
  ix->mask_lr = ((u64m >> (64 - k)) << 32) + ((u64m << 32) >> (64 - k));
  ix->mask_bp = u64m >> ((32 - k) * 2);
}

/* ------------------------------------------------------------------------------------------------ likelihood */

/* ref src/hdhistllh.hpp:51-69 */
void ko_llh_tables(uint32_t h, uint32_t k, uint32_t th, uint64_t* ck, uint64_t* hnk)
{
  uint64_t vc = 1;
  uint32_t nh = k - h;
  ck[0] = 1; hnk[0] = 0;
  for (uint32_t i = 0; i < k; ++i) ck[i + 1] = (ck[i] * (k - i)) / (i + 1);
  for (uint32_t i = 1; i <= th; ++i) { vc = (vc * (nh - i + 1)) / i; hnk[i] = ck[i] - vc; }
}

typedef struct { uint32_t h, k, th; uint64_t ck[65], hnk[KO_MAX_TH + 1]; const double* mc; double uc, rho; } llh_t;

/* ref src/hdhistllh.hpp:71-89, literal evaluation order */
static double llh_eval(const llh_t* f, double d)
{
  double sum = 0.0, lv_m = 0.0;
  double powdc = pow((1.0 - d), (double)f->k);
  double logdn = log(1.0 - d);
  double logdp = log(d) - logdn;
  logdn *= f->k;
  double dratio = d / (1.0 - d);
  for (uint32_t x = 0; x <= f->k; ++x) {
    if (x <= f->th) {
      sum -= (logdn + x * logdp) * f->mc[x];
      lv_m += (double)f->hnk[x] * powdc;
    } else {
      lv_m += powdc * (double)f->ck[x];
    }
    powdc *= dratio;
  }
  return sum - log(f->rho * lv_m + 1.0 - f->rho) * f->uc;
}

double ko_llh(uint32_t h, uint32_t k, uint32_t th, const double* hist, double uc, double rho, double d)
{
  llh_t f; f.h = h; f.k = k; f.th = th; f.mc = hist; f.uc = uc; f.rho = rho;
  ko_llh_tables(h, k, th, f.ck, f.hnk);
  return llh_eval(&f, d);
}

/* ref external/boost/libs/math/include/boost/math/tools/minima.hpp:23-138 with (min,max,bits)=(1e-10,0.5,16) as called
 * at src/query.cpp:430.  bits = min(digits<double>/2 = 26, 16) = 16; tolerance = ldexp(1.0, 1-16). */
static void brent(const llh_t* f, double* xo, double* fo, uint32_t* iters)
{
  double min = 1e-10, max = 0.5;
  const double tolerance = ldexp(1.0, 1 - 16);
  double x, w, v, u, delta, delta2, fu, fv, fw, fx, mid, fract1, fract2;
  static const double golden = 0.3819660f;
  x = w = v = max;
  fw = fv = fx = llh_eval(f, x);
  delta2 = delta = 0;
  uint32_t it = 0;
  for (;;) {
    mid = (min + max) / 2;
    fract1 = tolerance * fabs(x) + tolerance / 4;
    fract2 = 2 * fract1;
    if (fabs(x - mid) <= (fract2 - (max - min) / 2)) break;
    if (fabs(delta2) > fract1) {
      double r = (x - w) * (fx - fv);
      double q = (x - v) * (fx - fw);
      double p = (x - v) * q - (x - w) * r;
      q = 2 * (q - r);
      if (q > 0) p = -p;
      q = fabs(q);
      double td = delta2;
      delta2 = delta;
      if ((fabs(p) >= fabs(q * td / 2)) || (p <= q * (min - x)) || (p >= q * (max - x))) {
        delta2 = (x >= mid) ? min - x : max - x;
        delta = golden * delta2;
      } else {
        delta = p / q;
        u = x + delta;
        if (((u - min) < fract2) || ((max - u) < fract2)) delta = (mid - x) < 0 ? -fabs(fract1) : fabs(fract1);
      }
    } else {
      delta2 = (x >= mid) ? min - x : max - x;
      delta = golden * delta2;
    }
    u = (fabs(delta) >= fract1) ? (x + delta) : (delta > 0 ? (x + fabs(fract1)) : (x - fabs(fract1)));
    fu = llh_eval(f, u);
    ++it;
    if (fu <= fx) {
      if (u >= x) min = x; else max = x;
      v = w; w = x; x = u; fv = fw; fw = fx; fx = fu;
    } else {
      if (u < x) min = u; else max = u;
      if ((fu <= fw) || (w == x)) { v = w; w = u; fv = fw; fw = fu; }
      else if ((fu <= fv) || (v == x) || (v == w)) { v = u; fv = fu; }
    }
  }
  *xo = x; *fo = fx;
  if (iters) *iters = it;
}

void ko_brent(uint32_t h, uint32_t k, uint32_t th, const double* hist, double uc, double rho, double* d, double* v,
              uint32_t* iters)
{
  llh_t f; f.h = h; f.k = k; f.th = th; f.mc = hist; f.uc = uc; f.rho = rho;
  ko_llh_tables(h, k, th, f.ck, f.hnk);
  brent(&f, d, v, iters);
}

/* ------------------------------------------------------------------------------------------------ tree */

typedef struct { char** v; size_t n, cap; } toks_t;
static void tok_push(toks_t* t, const char* s, size_t len)
{
  if (t->n == t->cap) { t->cap = t->cap ? 2 * t->cap : 64; t->v = (char**)realloc(t->v, t->cap * sizeof(char*)); }
  char* c = (char*)malloc(len + 1); memcpy(c, s, len); c[len] = 0; t->v[t->n++] = c;
}

/* ref src/phytree.cpp:84-148 (Tree::split_nwk).  Returns 0 on success. */
static int split_nwk(const char* nwk_in, size_t n, toks_t* el, char* err, size_t errlen)
{
  if (n == 0) { seterr(err, errlen, "Given Newick tree seems to be empty?!?."); return -1; }
  if (nwk_in[n - 1] == '\n') n--;
  if (n == 0 || nwk_in[n - 1] != ';') { seterr(err, errlen, "Given Newick tree ends with a character other than ';'."); return -1; }
  char* buf = (char*)malloc(n + 2); size_t bl = 0;
  int is_quoted = 0, quote = 0, quote_p = 0, is_comment = 0;
  for (size_t i = 0; i < n; i++) {
    char c = nwk_in[i];
    if (is_comment) { is_comment = is_comment != (c == ']'); continue; }
    quote = (c == '\'' || c == '"');
    if (quote & quote_p) { is_quoted = 0; buf[bl++] = '\''; continue; }
    quote_p = quote;
    if (quote) { is_quoted = (is_quoted != quote); continue; }
    else if (is_quoted) { is_comment = is_comment != (c == '['); if (!is_comment) buf[bl++] = c; }
    else if (c == '(' || c == ')' || c == ':' || c == ',') {
      if (c != '(' && (i == 0 || nwk_in[i - 1] != '(')) { tok_push(el, buf, bl); bl = 0; }
      tok_push(el, &c, 1);
    } else {
      if (c == '[' || c == ']') { seterr(err, errlen, "Given Newick tree contains an unquoted label or length with '[' or ']'."); free(buf); return -1; }
      if (c == ';') {
        if (i == n - 1) break;
        seterr(err, errlen, "Given Newick tree contains an unexpected ';'."); free(buf); return -1;
      }
      if ((c == ' ' || c == '\n') && bl) { seterr(err, errlen, "Given Newick tree contains an unquoted label or length with ' ' or newline."); free(buf); return -1; }
      buf[bl++] = c;
    }
  }
  if (bl > 0) tok_push(el, buf, bl);
  free(buf);
  return 0;
}

typedef struct pnode { struct pnode* parent; struct pnode** ch; uint32_t nch, capch, se, card; int is_leaf; double blen; char* name; } pnode_t;
typedef struct { toks_t* el; size_t at; uint32_t nnodes; pnode_t** by_se; size_t cap; int fail; char* err; size_t errlen; } pctx_t;

static int tok_is(pctx_t* c, const char* s) { return c->at < c->el->n && strcmp(c->el->v[c->at], s) == 0; }
static void reg_node(pctx_t* c, pnode_t* nd)
{
  c->nnodes++; nd->se = c->nnodes;
  if (c->nnodes + 1 > c->cap) { c->cap = c->cap ? 2 * c->cap : 64; c->by_se = (pnode_t**)realloc(c->by_se, c->cap * sizeof(pnode_t*)); }
  c->by_se[c->nnodes] = nd;
}
static void name_blen(pctx_t* c, pnode_t* nd)
{ /* ref src/phytree.cpp:175-187,192-203: optional label, optional ":length" */
  nd->name = strdup(""); nd->blen = NAN;
  if (c->at >= c->el->n) return; /* the reference reads past the token vector here (unlabeled root); treated as no label */
  if (!tok_is(c, ",")) {
    if (!tok_is(c, ":")) { free(nd->name); nd->name = strdup(c->el->v[c->at]); c->at++; }
    if (tok_is(c, ":")) { nd->blen = (c->at + 1 < c->el->n) ? atof(c->el->v[c->at + 1]) : 0.0; c->at += 2; }
  }
}
/* ref src/phytree.cpp:150-215 (Node::parse): recursive descent, se assigned after the children (post-order) */
static void parse_node(pctx_t* c, pnode_t* nd)
{
  if (c->fail || c->at >= c->el->n) return;
  if (tok_is(c, "(")) {
    for (;;) {
      c->at++;
      pnode_t* ch = (pnode_t*)calloc(1, sizeof(pnode_t)); ch->is_leaf = 1; ch->blen = 0;
      parse_node(c, ch);
      ch->parent = nd; /* set_parent/add_children: ref src/phytree.hpp:95-116 */
      if (nd->nch == nd->capch) { nd->capch = nd->capch ? 2 * nd->capch : 4; nd->ch = (pnode_t**)realloc(nd->ch, nd->capch * sizeof(pnode_t*)); }
      nd->ch[nd->nch++] = ch; nd->card += ch->card; nd->is_leaf = 0;
      if (tok_is(c, ",")) continue; else break;
    }
    if (nd->nch == 1) { c->fail = 1; seterr(c->err, c->errlen, "A node has a single child in the backbone tree! Please suppress unifurcations."); return; }
    reg_node(c, nd);
    if (tok_is(c, ")")) { c->at++; if (tok_is(c, ")")) { if (!nd->name) nd->name = strdup(""); return; } }
    name_blen(c, nd);
  } else {
    name_blen(c, nd);
    nd->is_leaf = 1; nd->card = 1;
    reg_node(c, nd);
  }
}
static void free_pnode(pnode_t* nd) { for (uint32_t i = 0; i < nd->nch; ++i) free_pnode(nd->ch[i]); free(nd->ch); free(nd->name); free(nd); }

static int tree_from_newick(ko_tree_t* t, const char* nwk, size_t n, char* err, size_t errlen)
{
  toks_t el = {0};
  if (split_nwk(nwk, n, &el, err, errlen)) return -1;
  pctx_t c; memset(&c, 0, sizeof c); c.el = &el; c.err = err; c.errlen = errlen;
  pnode_t* root = (pnode_t*)calloc(1, sizeof(pnode_t)); root->is_leaf = 1;
  parse_node(&c, root);
  if (c.fail) { free_pnode(root); return -1; }
  uint32_t N = c.nnodes;
  t->nnodes = N; t->root = root->se;
  t->parent = (uint32_t*)calloc(N + 1, 4); t->nchildren = (uint32_t*)calloc(N + 1, 4); t->eff_nch = (uint32_t*)calloc(N + 1, 4);
  t->is_leaf = (uint8_t*)calloc(N + 1, 1); t->card = (uint32_t*)calloc(N + 1, 4); t->blen = (double*)calloc(N + 1, 8);
  t->name = (char**)calloc(N + 1, sizeof(char*)); t->children = (uint32_t**)calloc(N + 1, sizeof(uint32_t*));
  t->name[0] = strdup("");
  for (uint32_t se = 1; se <= N; ++se) {
    pnode_t* nd = c.by_se[se];
    t->parent[se] = nd->parent ? nd->parent->se : 0;
    t->nchildren[se] = nd->nch; t->eff_nch[se] = nd->nch; t->is_leaf[se] = (uint8_t)nd->is_leaf; t->card[se] = nd->card;
    t->blen[se] = nd->blen; t->name[se] = strdup(nd->name ? nd->name : "");
    t->children[se] = (uint32_t*)calloc(nd->nch ? nd->nch : 1, 4);
    for (uint32_t i = 0; i < nd->nch; ++i) t->children[se][i] = nd->ch[i]->se;
  }
  free_pnode(root); free(c.by_se);
  for (size_t i = 0; i < el.n; ++i) free(el.v[i]);
  free(el.v);
  return 0;
}

static void tree_free(ko_tree_t* t)
{
  if (!t->parent) return;
  for (uint32_t se = 0; se <= t->nnodes; ++se) { free(t->name[se]); free(t->children[se]); }
  free(t->parent); free(t->nchildren); free(t->eff_nch); free(t->is_leaf); free(t->card); free(t->blen); free(t->name); free(t->children);
  memset(t, 0, sizeof *t);
}

/* ------------------------------------------------------------------------------------------------ index loading */

/* ref src/phytree.cpp:217-253 (Node::generate_tree): a range of one name is a leaf; a longer range [first, last) gets two
 * children, the SECOND half [first + size/2, last) created before the first half, every branch length 1, internal nodes
 * unnamed.  Written as Newick (names quoted, a quote doubled) so that the post-order numbering of the parser reproduces se. */
static void gen_newick(char** names, size_t lo, size_t hi, char** out, size_t* len, size_t* cap)
{
#define KO_PUT(ch) do { if (*len + 2 > *cap) { *cap = *cap * 2 + 64; *out = (char*)realloc(*out, *cap); } (*out)[(*len)++] = (ch); } while (0)
  if (hi - lo == 1) {
    KO_PUT('\'');
    for (const char* c = names[lo]; *c; ++c) { if (*c == '\'') KO_PUT('\''); KO_PUT(*c); }
    KO_PUT('\'');
  } else {
    const size_t half = lo + (hi - lo) / 2;
    KO_PUT('(');
    gen_newick(names, half, hi, out, len, cap);
    KO_PUT(',');
    gen_newick(names, lo, half, out, len, cap);
    KO_PUT(')');
  }
  KO_PUT(':'); KO_PUT('1');
#undef KO_PUT
}

static char* newick_from_reflist(char* text, size_t n, size_t* out_len)
{
  size_t cnt = 0, capn = 16;
  char** names = (char**)malloc(capn * sizeof(char*));
  size_t at = 0;
  while (at < n) { /* std::getline: one name per line, a final line without newline counts */
    size_t e = at;
    while (e < n && text[e] != '\n') ++e;
    if (cnt == capn) { capn *= 2; names = (char**)realloc(names, capn * sizeof(char*)); }
    names[cnt] = (char*)malloc(e - at + 1); memcpy(names[cnt], text + at, e - at); names[cnt][e - at] = 0; ++cnt;
    at = e + 1;
  }
  char* out = NULL; size_t len = 0, cap = 0;
  if (cnt) { gen_newick(names, 0, cnt, &out, &len, &cap); out[len++] = ';'; out[len] = 0; }
  for (size_t i = 0; i < cnt; ++i) free(names[i]);
  free(names);
  *out_len = len;
  return out;
}

static char* slurp(const char* path, size_t* n)
{
  FILE* f = fopen(path, "rb"); if (!f) return NULL;
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  char* b = (char*)malloc((size_t)sz + 1);
  if (fread(b, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(b); return NULL; }
  fclose(f); b[sz] = 0; *n = (size_t)sz; return b;
}

static int cmpstr(const void* a, const void* b) { return strcmp(*(char* const*)a, *(char* const*)b); }

/* ref src/krepp.cpp:66-108 (TargetIndex::load_index), src/index.cpp:29-158, :188-201 */
ko_index_t* ko_index_load(const char* dir, char* err, size_t errlen)
{
  DIR* d = opendir(dir);
  if (!d) { seterr(err, errlen, "cannot open index directory %s", dir); return NULL; }
  char* sfx[64]; int ns = 0; struct dirent* e;
  while ((e = readdir(d))) {
    if (strncmp(e->d_name, "metadata-", 9) == 0 && !strchr(e->d_name, '.') && ns < 64) sfx[ns++] = strdup(e->d_name + 8);
  }
  closedir(d);
  if (!ns) { seterr(err, errlen, "no partial index found in %s", dir); return NULL; }
  qsort(sfx, ns, sizeof(char*), cmpstr);
  ko_index_t* ix = (ko_index_t*)calloc(1, sizeof *ix);
  ix->tables = (ko_table_t*)calloc(ns, sizeof(ko_table_t));
  char path[4096];
  for (int s = 0; s < ns; ++s) {
    ko_table_t* t = &ix->tables[ix->ntables];
    /* metadata: ref src/krepp.cpp:18-29 (writer), src/index.cpp:58-72 (reader) */
    snprintf(path, sizeof path, "%s/metadata%s", dir, sfx[s]);
    size_t n; unsigned char* md = (unsigned char*)slurp(path, &n);
    if (!md || n < 16) { seterr(err, errlen, "Failed to open %s", path); goto fail; }
    uint32_t k = md[0], w = md[1], h = md[2], m, r, nrows; uint8_t frac = md[11];
    memcpy(&m, md + 3, 4); memcpy(&r, md + 7, 4); memcpy(&nrows, md + 12, 4);
    if (n < 16 + k || h >= k || k > 32 || k - h > 16) { seterr(err, errlen, "bad metadata %s", path); free(md); goto fail; }
    if (ix->ntables == 0) {
      ix->k = k; ix->w = w; ix->h = h; ix->m = m;
      memcpy(ix->ppos, md + 16, h); memcpy(ix->npos, md + 16 + h, k - h);
      set_masks(ix);
      ix->res_table = (int32_t*)malloc(m * sizeof(int32_t)); ix->res_numer = (uint32_t*)calloc(m, 4);
      for (uint32_t i = 0; i < m; ++i) ix->res_table[i] = -1;
    } else if (k != ix->k || h != ix->h || m != ix->m || memcmp(ix->ppos, md + 16, h) || memcmp(ix->npos, md + 16 + h, k - h)) {
      seterr(err, errlen, "Partial libraries have incompatible hash functions!"); free(md); goto fail;
    }
    free(md);
    t->r = r; t->frac = frac;
    /* tree: ref src/index.cpp:29-49; without a tree file the balanced tree over reflist-* (generate_partial_tree, ref
     * src/index.cpp:3-27, Node::generate_tree src/phytree.cpp:217-253) is restated as the Newick text the parser then reads */
    if (!ix->have_tree) {
      snprintf(path, sizeof path, "%s/tree%s", dir, sfx[s]);
      char* nwk = slurp(path, &n);
      if (!nwk) {
        snprintf(path, sizeof path, "%s/reflist%s", dir, sfx[s]);
        char* rl = slurp(path, &n);
        if (!rl) { seterr(err, errlen, "Unable to open reference list file for an index without a tree."); goto fail; }
        nwk = newick_from_reflist(rl, n, &n); free(rl);
        if (!nwk) { seterr(err, errlen, "Unable to open reference list file for an index without a tree."); goto fail; }
      }
      int rc = tree_from_newick(&ix->tree, nwk, n, err, errlen); free(nwk);
      if (rc) goto fail;
      ix->have_tree = 1;
    }
    /* cmer + inc: ref src/table.cpp:65-75 */
    snprintf(path, sizeof path, "%s/cmer%s", dir, sfx[s]);
    char* b = slurp(path, &n);
    if (!b || n < 8) { seterr(err, errlen, "Failed to open %s", path); goto fail; }
    memcpy(&t->nkmers, b, 8);
    if (n < 8 + 8 * t->nkmers) { seterr(err, errlen, "Failed to read the k-mer vector of a partial index!"); free(b); goto fail; }
    t->cmer = (uint32_t*)malloc(8 * t->nkmers + 8); memcpy(t->cmer, b + 8, 8 * t->nkmers); free(b);
    snprintf(path, sizeof path, "%s/inc%s", dir, sfx[s]);
    b = slurp(path, &n);
    if (!b || n < 4) { seterr(err, errlen, "Failed to open %s", path); goto fail; }
    memcpy(&t->nrows, b, 4);
    if (n < 4 + 8ull * t->nrows) { seterr(err, errlen, "Failed to read the offset array of a partial index!"); free(b); goto fail; }
    t->inc = (uint64_t*)malloc(8ull * t->nrows + 8); memcpy(t->inc, b + 4, 8ull * t->nrows); free(b);
    /* crecord: ref src/record.cpp:203-211 */
    snprintf(path, sizeof path, "%s/crecord%s", dir, sfx[s]);
    b = slurp(path, &n);
    if (!b || n < 8) { seterr(err, errlen, "Failed to open %s", path); goto fail; }
    memcpy(&t->nnodes, b, 4); memcpy(&t->nsubsets, b + 4, 4);
    if (n < 8 + 8ull * t->nsubsets + 8ull * t->nnodes) { seterr(err, errlen, "Failed to read the color array of a partial index!"); free(b); goto fail; }
    t->pse = (uint32_t*)malloc(8ull * t->nsubsets + 8); memcpy(t->pse, b + 8, 8ull * t->nsubsets);
    t->rho = (double*)malloc(8ull * t->nnodes + 8); memcpy(t->rho, b + 8 + 8ull * t->nsubsets, 8ull * t->nnodes); free(b);
    /* residue registration: ref src/index.cpp:144-157 */
    if (frac) { for (uint32_t i = 0; i <= r && i < ix->m; ++i) { ix->res_table[i] = (int32_t)ix->ntables; ix->res_numer[i] = r + 1; } }
    else if (r < ix->m) { ix->res_table[r] = (int32_t)ix->ntables; ix->res_numer[r] = 1; }
    ix->ntables++;
  }
  { /* make_rho_partial: ref src/index.cpp:188-201, src/record.cpp:304-309 */
    uint32_t present = 0;
    for (uint32_t i = 0; i < ix->m; ++i) present += ix->res_table[i] >= 0;
    double ratio_m = (double)present / (double)ix->m;
    for (uint32_t t = 0; t < ix->ntables; ++t)
      for (uint32_t se = 0; se < ix->tables[t].nnodes; ++se) ix->tables[t].rho[se] *= ratio_m;
  }
  for (int s = 0; s < ns; ++s) free(sfx[s]);
  return ix;
fail:
  for (int s = 0; s < ns; ++s) free(sfx[s]);
  ko_index_free(ix);
  return NULL;
}

void ko_index_free(ko_index_t* ix)
{
  if (!ix) return;
  for (uint32_t t = 0; t < ix->ntables; ++t) { free(ix->tables[t].cmer); free(ix->tables[t].inc); free(ix->tables[t].pse); free(ix->tables[t].rho); }
  free(ix->tables); free(ix->res_table); free(ix->res_numer);
  tree_free(&ix->tree);
  free(ix);
}

ko_index_t* ko_geom_new(uint32_t k, uint32_t h, uint32_t m, uint32_t r, int frac, const uint8_t* ppos)
{
  ko_index_t* ix = (ko_index_t*)calloc(1, sizeof *ix);
  ix->k = k; ix->h = h; ix->m = m; ix->geom_r = r; ix->geom_frac = (uint32_t)frac;
  memcpy(ix->ppos, ppos, h);
  /* npos = complement of ppos in 0..k-1, ascending (ref src/lshf.cpp:138-146) */
  uint32_t j = 0;
  for (uint32_t i = 0; i < k; ++i) { int in = 0; for (uint32_t q = 0; q < h; ++q) in |= (ppos[q] == i); if (!in) ix->npos[j++] = (uint8_t)i; }
  set_masks(ix);
  return ix;
}

uint32_t ko_index_k(const ko_index_t* ix) { return ix->k; }
uint32_t ko_index_h(const ko_index_t* ix) { return ix->h; }
uint32_t ko_index_m(const ko_index_t* ix) { return ix->m; }
uint32_t ko_index_nnodes(const ko_index_t* ix) { return ix->tree.nnodes; }
uint64_t ko_index_mask_hash_bp(const ko_index_t* ix) { return ix->mask_hash_bp; }
uint64_t ko_index_mask_drop_lr(const ko_index_t* ix) { return ix->mask_drop_lr; }
int ko_index_is_leaf(const ko_index_t* ix, uint32_t se) { return se && se <= ix->tree.nnodes ? ix->tree.is_leaf[se] : 0; }
uint32_t ko_index_parent(const ko_index_t* ix, uint32_t se) { return se && se <= ix->tree.nnodes ? ix->tree.parent[se] : 0; }
double ko_index_blen(const ko_index_t* ix, uint32_t se) { return se && se <= ix->tree.nnodes ? ix->tree.blen[se] : NAN; }

/* ref src/phytree.hpp:133-144 (Node::get_name(false)): label, or to_string(se-1) when unlabeled */
const char* ko_index_node_name(const ko_index_t* ix, uint32_t se)
{
  static __thread char buf[32];
  if (!se || se > ix->tree.nnodes) return "";
  if (ix->tree.name[se][0]) return ix->tree.name[se];
  snprintf(buf, sizeof buf, "%u", se - 1);
  return buf;
}

/* ref src/index.hpp:27 (check_partial), src/index.cpp:160-168 (bucket_indices), src/table.hpp:121-136 */
static const ko_table_t* bucket_of(const ko_index_t* ix, uint32_t rix, uint64_t* b, uint64_t* e)
{
  uint32_t res = rix % ix->m, offset = rix / ix->m;
  if (ix->res_table[res] < 0) return NULL;
  const ko_table_t* t = &ix->tables[ix->res_table[res]];
  if (ix->res_numer[res] > 1) offset = offset * ix->res_numer[res] + res;
  *b = offset ? t->inc[offset - 1] : 0;
  *e = offset < t->nrows ? t->inc[offset] : t->nkmers;
  return t;
}
int ko_index_bucket(const ko_index_t* ix, uint32_t rix, uint64_t* begin, uint64_t* end) { return bucket_of(ix, rix, begin, end) != NULL; }

/* ------------------------------------------------------------------------------------------------ per-read query */

typedef struct {
  int used; uint32_t last_pos, last_hdist, hdist_min;
  double match_count, mismatch_count, nmers, rho, hist[KO_MAX_TH + 1];
} acc_t;

/* ref src/query.hpp:153-176 (Minfo::update_match) */
static void update_match(acc_t* a, uint32_t pos, uint32_t hd)
{
  if (a->last_hdist == 0xFFFFFFFFu || a->last_pos != pos) {
    a->match_count++; a->mismatch_count--; a->hist[hd]++; a->last_pos = pos; a->last_hdist = hd;
  } else if (a->last_hdist > hd) {
    a->hist[hd]++; a->hist[a->last_hdist]--; a->last_hdist = hd;
  }
  if (hd < a->hdist_min) a->hdist_min = hd;
}

typedef struct { uint32_t* v; size_t head, tail, cap; } queue_t;
static void q_push(queue_t* q, uint32_t x) { if (q->tail == q->cap) { q->cap = q->cap ? q->cap * 2 : 64; q->v = (uint32_t*)realloc(q->v, q->cap * 4); } q->v[q->tail++] = x; }

/* ref src/query.cpp:352-390 (IMers::add_matching_mer) */
static void add_matching_mer(const ko_index_t* ix, const ko_params_t* p, acc_t* acc, uint32_t* hdist_filt, uint32_t enmers,
                             uint32_t pos, uint32_t rix, uint32_t enc_lr, queue_t* q)
{
  uint64_t b = 0, e = 0;
  const ko_table_t* t = bucket_of(ix, rix, &b, &e);
  if (!t) return;
  const ko_tree_t* tr = &ix->tree;
  for (; b < e; ++b) {
    uint32_t hd = ko_popcount_lr32(t->cmer[2 * b] ^ enc_lr);
    if (hd > p->hdist_th) continue;
    if (hd < *hdist_filt) *hdist_filt = hd;
    q->head = q->tail = 0;
    q_push(q, t->cmer[2 * b + 1]);
    while (q->head < q->tail) {
      uint32_t se = q->v[q->head++];
      if (se <= tr->nnodes) {              /* Tree::check_node */
        if (se == 0) continue;             /* se_to_node[0] is the null sentinel */
        if (tr->is_leaf[se]) {
          acc_t* a = &acc[se];
          if (!a->used) { /* Minfo(hdist_th, enmers, rho): ref src/query.hpp:116-123 */
            memset(a, 0, sizeof *a); a->used = 1; a->nmers = enmers; a->mismatch_count = enmers; a->rho = t->rho[se];
            a->last_hdist = 0xFFFFFFFFu; a->hdist_min = 0xFFFFFFFFu;
          }
          update_match(a, pos, hd);
          continue;
        }
      }
      if (se >= t->nsubsets) continue; /* out of the reference's domain (it would read past se_to_pse) */
      q_push(q, t->pse[2 * se]); q_push(q, t->pse[2 * se + 1]);
    }
  }
}

static void minfo_from_acc(ko_minfo_t* m, const acc_t* a, uint32_t strand, uint32_t se)
{
  memset(m, 0, sizeof *m);
  m->strand = strand; m->leaf_se = se; m->hdist_min = a->hdist_min; m->match_count = a->match_count;
  m->mismatch_count = a->mismatch_count; m->rho = a->rho; m->nmers = a->nmers;
  memcpy(m->hist, a->hist, sizeof m->hist);
  m->d_llh = DBL_MAX; m->v_llh = NAN; /* ref src/query.hpp:225-226 */
}

static void optimize(const ko_index_t* ix, const ko_params_t* p, const llh_t* proto, const double* hist, double uc, double rho, double* d, double* v)
{ /* ref src/query.cpp:426-433 */
  llh_t f = *proto; f.mc = hist; f.uc = uc; f.rho = rho;
  brent(&f, d, v, NULL);
}
static double likelihood_ratio(const llh_t* proto, const ko_minfo_t* cl, double d)
{ /* ref src/query.cpp:420-424 */
  llh_t f = *proto; f.mc = cl->hist; f.uc = cl->mismatch_count; f.rho = cl->rho;
  return 2 * (llh_eval(&f, d) - cl->v_llh);
}

typedef struct { int used; uint32_t hdist_min, rmatch; double nmers, mismatch_count, match_count, rho, hist[KO_MAX_TH + 1], d_llh, v_llh, chisq, lwr; } anc_t;

void ko_query_read(const ko_index_t* ix, const ko_params_t* p, const char* seq, uint64_t len, ko_read_t* out)
{
  const ko_tree_t* tr = &ix->tree;
  uint32_t k = ix->k, N = tr->nnodes, th = p->hdist_th;
  memset(out, 0, sizeof *out);
  out->len = len; out->closest = -1;
  out->hdist_filt[0] = out->hdist_filt[1] = 0xFFFFFFFFu;
  acc_t* acc[2];
  acc[0] = (acc_t*)calloc(N + 1, sizeof(acc_t)); acc[1] = (acc_t*)calloc(N + 1, sizeof(acc_t));
  queue_t q = {0};
  size_t lcap = 0;
  /* ---- search_mers: ref src/query.cpp:40-94 (non-CANONICAL branch :82-91) */
  uint32_t enmers = (uint32_t)(len - k + 1);
  uint32_t i, l, onmers = 0;
  uint64_t bp = 0, lr = 0, rcbp;
  for (i = l = 0; i < len;) {
    unsigned c = nt4((unsigned char)seq[i]);
    if (c >= 4) { l = 0; i++; continue; }
    l++; i++;
    if (l < k) {
      /* compute_encoding is evaluated at l==k over the last k bases; keep a rolling copy so that equals it */
      bp = (bp << 2) + c; lr = ((lr << 1) & 0xFFFFFFFEFFFFFFFEull) + (c & 1) + ((uint64_t)(c >> 1) << 32);
      continue;
    }
    /* ref src/common.hpp:225-243: appending base c */
    bp = (bp << 2) + c; lr = ((lr << 1) & 0xFFFFFFFEFFFFFFFEull) + (c & 1) + ((uint64_t)(c >> 1) << 32);
    bp &= ix->mask_bp; lr &= ix->mask_lr;
    rcbp = ko_revcomp_bp64(bp, k);
    onmers++;
    for (uint32_t st = 0; st < 2; ++st) {
      uint64_t ebp = st ? rcbp : bp, elr = st ? ko_conv_bp64_lr64(rcbp) : lr;
      uint32_t rix = (uint32_t)ko_pext64(ebp, ix->mask_hash_bp);       /* ref src/lshf.cpp:62 */
      if (ix->res_table[rix % ix->m] < 0) continue;                    /* ref src/index.hpp:27 */
      uint32_t enc = (uint32_t)ko_pext64(elr, ix->mask_drop_lr);       /* ref src/lshf.cpp:64-69 */
      uint32_t pos = st ? (uint32_t)(len - i) : (i - k);
      if (p->want_lookups) {
        if (out->n_lookups == lcap) { lcap = lcap ? 2 * lcap : 256; out->lookups = (ko_lookup_t*)realloc(out->lookups, lcap * sizeof(ko_lookup_t)); }
        ko_lookup_t L = {st, pos, rix, enc}; out->lookups[out->n_lookups++] = L;
      }
      add_matching_mer(ix, p, acc[st], &out->hdist_filt[st], enmers, pos, rix, enc, &q);
      out->wn[st]++;
    }
  }
  /* a run shorter than k that was reset by a non-ACGT base must not leak into the next run */
  /* (handled above: l restarts at 0 and bp/lr are masked to the last k bases once l reaches k) */
  out->onmers = onmers;
  /* ---- collect per-(strand, leaf) records */
  uint32_t nm = 0;
  for (uint32_t st = 0; st < 2; ++st) for (uint32_t se = 1; se <= N; ++se) nm += acc[st][se].used;
  out->minfo = (ko_minfo_t*)calloc(nm ? nm : 1, sizeof(ko_minfo_t));
  int32_t* mix[2]; mix[0] = (int32_t*)malloc((N + 1) * 4); mix[1] = (int32_t*)malloc((N + 1) * 4);
  for (uint32_t st = 0; st < 2; ++st) for (uint32_t se = 0; se <= N; ++se) {
    mix[st][se] = -1;
    if (se && acc[st][se].used) { mix[st][se] = (int32_t)out->n_minfo; minfo_from_acc(&out->minfo[out->n_minfo++], &acc[st][se], st, se); }
  }
  /* ---- summarize_matches: ref src/query.cpp:96-139, visiting order fixed:
   *      forward leaves by ascending se, then reverse leaves by ascending se; `<=` kept, so the last tie wins. */
  llh_t proto; proto.h = ix->h; proto.k = k; proto.th = th; proto.mc = NULL; proto.uc = 0; proto.rho = 0;
  ko_llh_tables(ix->h, k, th, proto.ck, proto.hnk);
  uint32_t filt[2] = {2 * out->hdist_filt[0] + 1, 2 * out->hdist_filt[1] + 1};
  int32_t* selof = (int32_t*)malloc((N + 1) * 4); /* leaf se -> minfo index chosen for node_to_minfo */
  for (uint32_t se = 0; se <= N; ++se) selof[se] = -1;
  int32_t cl = -1; double cl_d = DBL_MAX;
  for (uint32_t st = 0; st < 2; ++st) {
    for (uint32_t se = 1; se <= N; ++se) {
      int32_t j = mix[st][se]; if (j < 0) continue;
      ko_minfo_t* mi = &out->minfo[j];
      mi->mismatch_count = onmers - mi->match_count;
      if (mi->hdist_min > filt[st]) continue;
      optimize(ix, p, &proto, mi->hist, mi->mismatch_count, mi->rho, &mi->d_llh, &mi->v_llh);
      mi->solved = 1;
      if (mi->d_llh <= cl_d) { cl = j; cl_d = mi->d_llh; }
      selof[se] = j;
      if (st == 1 && mix[0][se] >= 0) {
        const ko_minfo_t* mo = &out->minfo[mix[0][se]];
        if ((mi->d_llh > mo->d_llh) || ((mi->d_llh == mo->d_llh) && (mi->match_count < mo->match_count))) selof[se] = mix[0][se];
      }
    }
  }
  if (cl >= 0) selof[out->minfo[cl].leaf_se] = cl;
  uint32_t ns = 0;
  for (uint32_t se = 1; se <= N; ++se) ns += selof[se] >= 0;
  out->sel = (ko_sel_t*)calloc(ns ? ns : 1, sizeof(ko_sel_t));
  for (uint32_t se = 1; se <= N; ++se) {
    if (selof[se] < 0) continue;
    const ko_minfo_t* mi = &out->minfo[selof[se]];
    ko_sel_t* s = &out->sel[out->n_sel];
    s->leaf_se = se; s->strand = mi->strand; s->minfo_ix = (uint32_t)selof[se]; s->d_llh = mi->d_llh; s->v_llh = mi->v_llh;
    s->is_closest = (selof[se] == cl);
    s->chisq = likelihood_ratio(&proto, &out->minfo[cl], mi->d_llh);
    if (s->is_closest) out->closest = (int32_t)out->n_sel;
    out->n_sel++;
  }
  /* ---- report_placement: ref src/query.cpp:218-333 (multi, not summarize), ancestors visited by ascending se */
  if (p->want_place && out->n_sel) {
    const ko_minfo_t* mcl = &out->minfo[cl];
    double leq = 0; for (uint32_t x = 0; x <= p->tau && x <= th; ++x) leq += mcl->hist[x]; /* get_leq_tau: ref src/query.hpp:189-196 */
    if (p->no_filter || leq > 1.0) {
      out->place = (ko_place_t*)calloc(N + 1, sizeof(ko_place_t));
      if (out->n_sel == 1) {
        ko_place_t* pl = &out->place[out->n_place++];
        uint32_t se = mcl->leaf_se;
        double mid = isnan(tr->blen[se]) ? 0 : tr->blen[se] / 2.0;
        pl->se = se; pl->edge = se - 1; pl->d_llh = mcl->d_llh; pl->v_llh = mcl->v_llh; pl->chisq = 0; pl->lwr = 1;
        pl->pendant = -0.75 * log(1 - 4.0 / 3.0 * mcl->d_llh) - mid; pl->distal = mid;
      } else {
        anc_t* pp = (anc_t*)calloc(N + 1, sizeof(anc_t));
        for (uint32_t si = 0; si < out->n_sel; ++si) {
          const ko_minfo_t* mi = &out->minfo[out->sel[si].minfo_ix];
          uint32_t se = mi->leaf_se;
          anc_t* a = &pp[se]; a->used = 1; a->hdist_min = mi->hdist_min; a->nmers = mi->nmers; a->mismatch_count = mi->mismatch_count;
          a->match_count = mi->match_count; a->rho = mi->rho; memcpy(a->hist, mi->hist, sizeof a->hist); a->d_llh = mi->d_llh; a->v_llh = mi->v_llh;
          double denom = 1.0;
          for (uint32_t par = tr->parent[se]; par; par = tr->parent[par]) {
            denom /= tr->eff_nch[par]; /* check_taxon() is false without a lineage file: ref src/query.cpp:254-259 */
            anc_t* g = &pp[par];
            if (!g->used) { memset(g, 0, sizeof *g); g->used = 2; g->hdist_min = 0xFFFFFFFFu; g->d_llh = DBL_MAX; g->v_llh = NAN; }
            /* Minfo::add: ref src/query.hpp:139-152 */
            g->mismatch_count = g->nmers ? g->mismatch_count : mi->nmers;
            g->match_count += mi->match_count * denom;
            g->mismatch_count -= mi->match_count * denom;
            for (uint32_t x = 0; x <= th; ++x) g->hist[x] = g->hist[x] + mi->hist[x] * denom;
            if (mi->hdist_min < g->hdist_min) g->hdist_min = mi->hdist_min;
            if (mi->nmers > g->nmers) g->nmers = mi->nmers;
            if (mi->rho > g->rho) g->rho = mi->rho;
            g->rmatch++;
          }
        }
        double total = 0;
        for (uint32_t se = 1; se <= N; ++se) {
          anc_t* a = &pp[se]; if (!a->used) continue;
          if (tr->nchildren[se] != tr->eff_nch[se] || tr->nchildren[se] == 1) continue;
          double lq = 0; for (uint32_t x = 0; x <= p->tau && x <= th; ++x) lq += a->hist[x];
          if (!(p->no_filter || lq > 1.0)) continue;
          if (!tr->is_leaf[se]) optimize(ix, p, &proto, a->hist, a->mismatch_count, a->rho, &a->d_llh, &a->v_llh);
          a->chisq = likelihood_ratio(&proto, mcl, a->d_llh);
          if ((a->chisq < p->chisq) && tr->parent[se]) {
            ko_place_t* pl = &out->place[out->n_place++];
            double mid = isnan(tr->blen[se]) ? 0 : tr->blen[se] / 2.0;
            pl->se = se; pl->edge = se - 1; pl->d_llh = a->d_llh; pl->v_llh = a->v_llh; pl->chisq = a->chisq;
            pl->lwr = exp(-a->chisq / 2); total = total + pl->lwr;
            pl->pendant = -0.75 * log(1 - 4.0 / 3.0 * a->d_llh) - mid; pl->distal = mid;
          }
        }
        for (uint32_t j = 0; j < out->n_place; ++j) out->place[j].lwr = out->place[j].lwr / total;
        free(pp);
      }
    }
  }
  free(selof); free(mix[0]); free(mix[1]); free(acc[0]); free(acc[1]); free(q.v);
}

void ko_read_free(ko_read_t* r) { free(r->lookups); free(r->minfo); free(r->sel); free(r->place); memset(r, 0, sizeof *r); }

/* ------------------------------------------------------------------------------------------------ text I/O */

/* kseq framing: ref src/kseq.h:177-216.  name = header up to first whitespace; sequence = every isgraph() byte up to
 * the next '>', '@' or '+'; FASTQ quality is skipped by length. */
int64_t ko_parse_reads(const char* t, size_t n, ko_seqrec_t** out)
{
  size_t cap = 1024, cnt = 0, i = 0;
  ko_seqrec_t* v = (ko_seqrec_t*)malloc(cap * sizeof *v);
  int last = 0;
  for (;;) {
    if (!last) { /* jump to the next header line */
      while (i < n && t[i] != '>' && t[i] != '@') i++;
      if (i >= n) break;
      last = t[i++];
    }
    if (i >= n) break; /* ks_getuntil returns -1 at end of file */
    size_t s = i;
    while (i < n && !isspace((unsigned char)t[i])) i++;
    char* name = strndup(t + s, i - s);
    int c = i < n ? (unsigned char)t[i++] : 0;
    if (c != '\n') { while (i < n && t[i] != '\n') i++; if (i < n) i++; }
    size_t scap = 256, sl = 0; char* sq = (char*)malloc(scap);
    c = -1;
    while (i < n) {
      int ch = (unsigned char)t[i++];
      if (ch == '>' || ch == '+' || ch == '@') { c = ch; break; }
      if (isgraph(ch)) { if (sl + 2 > scap) { scap *= 2; sq = (char*)realloc(sq, scap); } sq[sl++] = (char)ch; }
    }
    sq[sl] = 0;
    if (c == '>' || c == '@') last = c;
    int ok = 1;
    if (c == '+') {
      while (i < n && t[i] != '\n') i++;
      if (i >= n) ok = 0; else i++;
      size_t ql = 0;
      while (ok && i < n) { int qc = (unsigned char)t[i++]; if (!(ql < sl)) break; if (qc >= 33 && qc <= 127) ql++; }
      last = 0;
      if (ok && ql != sl) ok = 0;
    }
    if (!ok) { free(name); free(sq); break; } /* kseq_read returns -2: the reader loop stops (ref src/rqseq.cpp:189) */
    if (cnt == cap) { cap *= 2; v = (ko_seqrec_t*)realloc(v, cap * sizeof *v); }
    v[cnt].name = name; v[cnt].seq = sq; v[cnt].len = sl; cnt++;
  }
  *out = v;
  return (int64_t)cnt;
}
void ko_free_reads(ko_seqrec_t* r, int64_t n) { for (int64_t i = 0; i < n; ++i) { free(r[i].name); free(r[i].seq); } free(r); }

typedef struct { char* s; size_t n, cap; } sbuf_t;
static void sb_printf(sbuf_t* b, const char* fmt, ...)
{
  va_list ap; va_start(ap, fmt); int need = vsnprintf(NULL, 0, fmt, ap); va_end(ap);
  if (b->n + (size_t)need + 1 > b->cap) { b->cap = (b->n + (size_t)need + 1) * 2; b->s = (char*)realloc(b->s, b->cap); }
  va_start(ap, fmt); vsnprintf(b->s + b->n, (size_t)need + 1, fmt, ap); va_end(ap);
  b->n += (size_t)need;
}

/* ref src/query.cpp:158-196 (report_distances; multi / no_filter / dist_max), DISTANCE_FIELDS src/query.hpp:210 */
static void report_distances(const ko_index_t* ix, const ko_params_t* p, const char* name, const ko_read_t* r, sbuf_t* b)
{
  int has_max = !isnan(p->dist_max);
  if (r->n_sel == 0 || (has_max && r->sel[r->closest].d_llh > p->dist_max)) { sb_printf(b, "%s\tNA\tNaN\n", name); return; }
  if (p->multi) {
    for (uint32_t i = 0; i < r->n_sel; ++i) {
      const ko_sel_t* s = &r->sel[i];
      if (!p->no_filter && !(s->chisq < p->chisq)) continue;
      if (has_max && !(s->d_llh < p->dist_max)) continue;
      sb_printf(b, "%s\t%s\t%.5f\n", name, ko_index_node_name(ix, s->leaf_se), s->d_llh);
    }
  } else {
    const ko_sel_t* s = &r->sel[r->closest];
    sb_printf(b, "%s\t%s\t%.5f\n", name, ko_index_node_name(ix, s->leaf_se), s->d_llh);
  }
}

char* ko_dist_tsv(const ko_index_t* ix, const ko_params_t* p, const ko_seqrec_t* recs, int64_t n, int nthreads)
{
  sbuf_t* parts = (sbuf_t*)calloc((size_t)(n ? n : 1), sizeof(sbuf_t));
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
  for (int64_t i = 0; i < n; ++i) {
    ko_read_t r;
    ko_query_read(ix, p, recs[i].seq, recs[i].len, &r);
    report_distances(ix, p, recs[i].name, &r, &parts[i]);
    ko_read_free(&r);
  }
  sbuf_t all = {0};
  sb_printf(&all, "%s", "");
  for (int64_t i = 0; i < n; ++i) { if (parts[i].n) sb_printf(&all, "%s", parts[i].s); free(parts[i].s); }
  free(parts);
  return all.s;
}

/* ref src/phytree.cpp:47-64 (stream_nwk_jplace), src/phytree.hpp:145-152 (stream_nwk_entry), numbers std::fixed 5 */
static void nwk_jplace(const ko_tree_t* t, uint32_t se, sbuf_t* b)
{
  if (!t->is_leaf[se]) {
    sb_printf(b, "(");
    for (uint32_t i = 0; i < t->nchildren[se]; ++i) { nwk_jplace(t, t->children[se][i], b); if (i + 1 < t->nchildren[se]) sb_printf(b, ","); }
    sb_printf(b, ")");
  }
  if (isnan(t->blen[se])) sb_printf(b, "%s", t->name[se]); else sb_printf(b, "%s:%.5f", t->name[se], t->blen[se]);
  sb_printf(b, "{%u}", se - 1);
  if (se == t->root) sb_printf(b, ";");
}
char* ko_index_jplace_tree(const ko_index_t* ix) { sbuf_t b = {0}; nwk_jplace(&ix->tree, ix->tree.root, &b); return b.s; }

/* ------------------------------------------------------------------------------------------------ index side (a17) */

static void mers_push(uint64_t** out, uint64_t* n, uint64_t* cap, uint64_t v)
{
  if (*n == *cap) { *cap = *cap ? *cap * 2 : 1024; *out = (uint64_t*)realloc(*out, *cap * 8); }
  (*out)[(*n)++] = v;
}

/* ref src/rqseq.cpp:51-144 (RSeq::extract_mers) with sdust off (its default): ring of w-k+1 slots, zero-initialised per
 * sequence, indexed by kix % ldiff and NOT reset by non-ACGT bases; an emit happens at every window end whose valid run
 * is >= w, and also at i == len (end-of-sequence quirk, :112-116).  set_curr_seq (src/rqseq.hpp:80-86) skips len < w. */
/* hll::HyperLogLog with b = 12 (ref src/hyperloglog.hpp:58-140): add() on the low 32 bits of the hash, estimate() */
#define KO_HLL_REGS 4096
static void hll_add(uint8_t* M, uint64_t z)
{
  const uint32_t hash = (uint32_t)z, index = hash >> 20, x = hash << 12;
  const uint32_t lz = x ? (uint32_t)__builtin_clz(x) : 32u; /* the reference's __builtin_clz(0) is undefined; 32 = lzcnt */
  const uint8_t rank = (uint8_t)((lz < 20 ? lz : 20) + 1);
  if (rank > M[index]) M[index] = rank;
}
static double hll_estimate(const uint8_t* M)
{
  const double m = (double)KO_HLL_REGS, alpha_mm = (0.7213 / (1.0 + 1.079 / m)) * m * m;
  double sum = 0.0;
  for (uint32_t i = 0; i < KO_HLL_REGS; ++i) sum += 1.0 / (double)(1 << M[i]);
  double e = alpha_mm / sum;
  if (e <= 2.5 * m) {
    uint32_t zeros = 0;
    for (uint32_t i = 0; i < KO_HLL_REGS; ++i) zeros += M[i] == 0;
    if (zeros) e = m * log(m / (double)zeros);
  } else if (e > (1.0 / 30.0) * 4294967296.0) e = -4294967296.0 * log(1.0 - e / 4294967296.0);
  return e;
}

static void extract_mers_est(const ko_index_t* g, const char* seq, uint64_t len, uint32_t w, uint64_t** out, uint64_t* n, uint64_t* cap, double* est);
void ko_extract_mers(const ko_index_t* g, const char* seq, uint64_t len, uint32_t w, uint64_t** out, uint64_t* n, uint64_t* cap)
{
  extract_mers_est(g, seq, len, w, out, n, cap, NULL);
}
/* the same walk with RSeq's two counters: est[0] += distinct valid k-mers, est[1] += distinct minimizers (ref src/rqseq.cpp:63-64,107-108,117,142-143) */
void ko_extract_mers_rho(const ko_index_t* g, const char* seq, uint64_t len, uint32_t w, uint64_t** out, uint64_t* n, uint64_t* cap, double* est)
{
  extract_mers_est(g, seq, len, w, out, n, cap, est);
}
static void extract_mers_est(const ko_index_t* g, const char* seq, uint64_t len, uint32_t w, uint64_t** out, uint64_t* n, uint64_t* cap, double* est)
{
  uint8_t* c1 = est ? (uint8_t*)calloc(2 * KO_HLL_REGS, 1) : NULL;
  uint8_t* c2 = c1 ? c1 + KO_HLL_REGS : NULL;
  uint32_t k = g->k, m = g->m, r = g->geom_r; int frac = (int)g->geom_frac;
  uint32_t ldiff;
  if (w > k) ldiff = w - k + 1; else { ldiff = 1; w = k; }
  if (len < w) { free(c1); return; }
  typedef struct { uint64_t x, y, z; } hm_t;
  hm_t* win = (hm_t*)calloc(ldiff, sizeof(hm_t));
  uint64_t kix = 0, bp = 0, lr = 0;
  uint64_t i, l;
  for (i = l = 0; i < len;) {
    unsigned c = nt4((unsigned char)seq[i]);
    if (c >= 4) { l = 0; i++; continue; }
    l++; i++;
    bp = (bp << 2) + c; lr = ((lr << 1) & 0xFFFFFFFEFFFFFFFEull) + (c & 1) + ((uint64_t)(c >> 1) << 32);
    if (l < k) continue;
    uint64_t x = bp & g->mask_bp, y = lr & g->mask_lr;
    hm_t cur = {x, y, ko_xur64_hash(x)};
    win[kix % ldiff] = cur; kix++;
    if (c1) hll_add(c1, cur.z);
    if ((l < w) && (i != len)) continue;
    hm_t mn = win[0];
    for (uint32_t j = 1; j < ldiff; ++j) if (win[j].z < mn.z) mn = win[j]; /* std::min_element: first minimum */
    if (c2) hll_add(c2, mn.z);
    uint32_t rix = (uint32_t)ko_pext64(mn.x, g->mask_hash_bp), res = rix % m;
    if (frac ? res <= r : res == r) {
      rix = frac ? rix / m * (r + 1) + res : rix / m;
      mers_push(out, n, cap, ((uint64_t)rix << 32) | (uint32_t)ko_pext64(mn.y, g->mask_drop_lr));
    }
  }
  free(win);
  if (est) { est[0] += hll_estimate(c1); est[1] += hll_estimate(c2); }
  free(c1);
}

/* ------------------------------------------------------------------------------------------------ seek (a sketch of one genome) */

struct ko_sketch {
  ko_index_t geom;  /* k, w, h, m, positions and masks; no tables */
  uint32_t r, frac, nrows;
  uint64_t nkmers;
  uint32_t* enc;    /* nkmers residual encodings (ref src/table.hpp SFlatHT::enc_v) */
  uint64_t* inc;    /* nrows cumulative ends */
  double rho;       /* after make_rho_partial */
};

/* ref src/sketch.cpp:3-23 Sketch::load_full_sketch (SFlatHT::load src/table.cpp:23-33, then the metadata, then rho) and
 * :25-32 make_rho_partial */
ko_sketch_t* ko_sketch_load(const char* path, char* err, size_t errlen)
{
  FILE* f = fopen(path, "rb");
  if (!f) { seterr(err, errlen, "Failed to open %s", path); return NULL; }
  ko_sketch_t* s = (ko_sketch_t*)calloc(1, sizeof *s);
  int ok = fread(&s->nkmers, 8, 1, f) == 1;
  if (ok) { s->enc = (uint32_t*)malloc((s->nkmers ? s->nkmers : 1) * 4); ok = fread(s->enc, 4, s->nkmers, f) == s->nkmers; }
  uint32_t nrows = 0;
  ok = ok && fread(&nrows, 4, 1, f) == 1;
  if (ok) { s->inc = (uint64_t*)malloc((nrows ? nrows : 1) * 8ull); ok = fread(s->inc, 8, nrows, f) == nrows; }
  uint8_t k = 0, w = 0, h = 0, frac = 0;
  uint32_t m = 0, r = 0, nrows2 = 0;
  ok = ok && fread(&k, 1, 1, f) == 1 && fread(&w, 1, 1, f) == 1 && fread(&h, 1, 1, f) == 1 && fread(&m, 4, 1, f) == 1 && fread(&r, 4, 1, f) == 1 &&
       fread(&frac, 1, 1, f) == 1 && fread(&nrows2, 4, 1, f) == 1;
  ok = ok && k && k <= 32 && h && h < k && m;
  if (ok) ok = fread(s->geom.ppos, 1, h, f) == h && fread(s->geom.npos, 1, (size_t)(k - h), f) == (size_t)(k - h) && fread(&s->rho, 8, 1, f) == 1;
  fclose(f);
  if (!ok) { seterr(err, errlen, "Failed to read the sketch file!"); ko_sketch_free(s); return NULL; }
  s->geom.k = k; s->geom.w = w; s->geom.h = h; s->geom.m = m;
  s->r = r; s->frac = frac; s->nrows = nrows;
  set_masks(&s->geom);
  s->rho *= frac ? ((double)r + 1.0) / (double)m : 1.0 / (double)m;
  return s;
}

void ko_sketch_free(ko_sketch_t* s) { if (s) { free(s->enc); free(s->inc); free(s); } }
uint32_t ko_sketch_k(const ko_sketch_t* s) { return s->geom.k; }
double ko_sketch_rho(const ko_sketch_t* s) { return s->rho; }

/* ref src/seek.cpp:22-53 seek_sequences (one sequence), :55-101 search_mers (non-CANONICAL), :103-120 add_matching_mer,
 * :121-127 optimize_likelihood; bucket addressing src/sketch.cpp:34-39, residue test src/sketch.hpp:18-22 */
void ko_seek_read(const ko_sketch_t* s, uint32_t th, const char* seq, uint64_t len, ko_seek_t* out)
{
  const ko_index_t* g = &s->geom;
  const uint32_t k = g->k;
  memset(out, 0, sizeof *out);
  out->dist = NAN; out->d[0] = out->d[1] = NAN; out->v[0] = out->v[1] = NAN;
  uint64_t i, l, onmers = 0;
  uint64_t bp = 0, lr = 0, rcbp;
  for (i = l = 0; i < len;) {
    unsigned c = nt4((unsigned char)seq[i]);
    if (c >= 4) { l = 0; i++; continue; }
    l++; i++;
    bp = (bp << 2) + c; lr = ((lr << 1) & 0xFFFFFFFEFFFFFFFEull) + (c & 1) + ((uint64_t)(c >> 1) << 32);
    if (l < k) continue;
    bp &= g->mask_bp; lr &= g->mask_lr;
    rcbp = ko_revcomp_bp64(bp, k);
    onmers++;
    for (uint32_t st = 0; st < 2; ++st) {
      const uint64_t ebp = st ? rcbp : bp, elr = st ? ko_conv_bp64_lr64(rcbp) : lr;
      const uint32_t rix = (uint32_t)ko_pext64(ebp, g->mask_hash_bp), res = rix % g->m;
      if (!((s->frac && res <= s->r) || res == s->r)) continue;
      const uint32_t enc = (uint32_t)ko_pext64(elr, g->mask_drop_lr);
      const uint64_t off = s->frac ? (uint64_t)(rix / g->m) * (s->r + 1) + res : rix / g->m;
      if (off >= s->nrows) continue; /* the reference would read past inc_v here; a valid sketch has the row */
      uint32_t hmin = th + 1;
      for (uint64_t e = off ? s->inc[off - 1] : 0; e < s->inc[off]; ++e) { const uint32_t hd = ko_popcount_lr32(s->enc[e] ^ enc); if (hd < hmin) hmin = hd; }
      if (hmin <= th) { out->match[st] += 1; out->hist[st][hmin] += 1; }
    }
  }
  out->onmers = onmers;
  if (out->match[0] + out->match[1] == 0) return;
  out->found = 1;
  for (uint32_t st = 0; st < 2; ++st) ko_brent(g->h, k, th, out->hist[st], (double)onmers - out->match[st], s->rho, &out->d[st], &out->v[st], NULL);
  out->dist = out->d[0] < out->d[1] ? out->d[0] : out->d[1];
}
