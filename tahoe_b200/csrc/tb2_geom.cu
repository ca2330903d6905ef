// tb2_geom.cu -- TahoeII text geometry (.geom) reader for meshes at BASELINE scale.  Host code only.
//
// The reference reads these files token by token through ifstreamT operator>> (toolbox/src/dataio/input/TahoeInputT.cpp,
// database/ModelFileT.cpp: GetDimensions / GetCoordinates / GetElementSet / GetNodeSet / GetSideSet); a 64M-element mesh is
// ~8 GB of text and minutes of that.  Here the file is read in one piece and the two bulk sections (*elements, *nodes -- inline or
// in the external files the main file names) are parsed by all host threads: chunks cut at line ends, one pass to count the
// numeric tokens per chunk, a prefix sum for each chunk's first (record, field), one pass to convert.  The small sections
// (dimensions, node sets, side sets) are parsed serially.  Numbers are converted with std::from_chars (correctly rounded, so
// coordinates are the same doubles the reference's stream extraction yields).  Layout of the file as in
// benchmark_XML/level.0/geometry/cube.1.geom and the generator level.5/explicit_benchmark/generate_3d_mesh.py:17-127.
#include "tb2_internal.h"

#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace tb2;

struct tb2_geom {
    int64_t nn = 0;
    int nsd = 0;
    struct Block { int id; int64_t nel; int nen; std::vector<int32_t> conn; };
    struct NodeSet { int id; std::vector<int32_t> nodes; };
    struct SideSet { int id, block; std::vector<int32_t> sides; };
    std::vector<Block> blocks;
    std::vector<NodeSet> nodesets;
    std::vector<SideSet> sidesets;
    std::vector<double> X;
};

namespace {

// TB2_GEOM_TIMING=1 prints the phases of tb2_geom_open to stderr
struct PhaseTimer {
    bool on = getenv("TB2_GEOM_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char* what)
    {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "tb2_geom: %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

struct Cursor {
    const char* p;
    const char* end;
};
inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
// next token [b, e); skips white space and '#' comments; false at the end
inline bool next_token(Cursor& c, const char*& b, const char*& e)
{
    const char* p = c.p;
    for (;;) {
        while (p < c.end && is_space(*p)) p++;
        if (p < c.end && *p == '#') {
            while (p < c.end && *p != '\n') p++;
            continue;
        }
        break;
    }
    if (p >= c.end) {
        c.p = p;
        return false;
    }
    b = p;
    while (p < c.end && !is_space(*p) && *p != '#') p++;
    e = p;
    c.p = p;
    return true;
}
inline bool to_i64(const char* b, const char* e, int64_t& v)
{
    if (b < e && *b == '+') b++;
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e;
}
inline bool to_f64(const char* b, const char* e, double& v)
{
    if (b < e && *b == '+') b++;
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e;
}
bool read_file(const std::string& path, std::vector<char>& buf)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    const size_t got = n ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

// `count` records of `per_rec` numeric tokens starting at c.p; field(rec, k, token) converts one token.  Returns false on a
// malformed or short section.  On success c.p is left after the last token consumed.
template <class Field>
bool parse_records(Cursor& c, int64_t count, int per_rec, Field field)
{
    const int64_t want = count * per_rec;
    if (want == 0) return true;
    // the section ends at the next '*' keyword (or the end of the buffer)
    const char* stop = (const char*)memchr(c.p, '*', (size_t)(c.end - c.p));
    if (!stop) stop = c.end;
    const size_t bytes = (size_t)(stop - c.p);
    int nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 32) nthreads = 32;
    if (bytes < (1u << 20)) nthreads = 1;
    std::vector<const char*> cut(nthreads + 1);
    cut[0] = c.p;
    cut[nthreads] = stop;
    for (int t = 1; t < nthreads; t++) {
        const char* q = c.p + bytes * (size_t)t / (size_t)nthreads;
        if (q < cut[t - 1]) q = cut[t - 1];
        while (q < stop && *q != '\n') q++; // a chunk starts after a line end: comments never straddle chunks
        cut[t] = q;
    }
    std::vector<int64_t> ntok(nthreads, 0);
    auto count_chunk = [&](int t) {
        Cursor k{cut[t], cut[t + 1]};
        const char *b, *e;
        int64_t n = 0;
        while (next_token(k, b, e)) n++;
        ntok[t] = n;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(count_chunk, t);
    count_chunk(0);
    for (auto& th : pool) th.join();
    pool.clear();
    std::vector<int64_t> first(nthreads + 1, 0);
    for (int t = 0; t < nthreads; t++) first[t + 1] = first[t] + ntok[t];
    if (first[nthreads] < want) return false; // short section (or a '*' inside a comment cut it: not worth a slow path)
    std::vector<char> ok(nthreads, 1);
    std::vector<const char*> last(nthreads, nullptr);
    auto parse_chunk = [&](int t) {
        Cursor k{cut[t], cut[t + 1]};
        const char *b, *e;
        int64_t idx = first[t];
        int64_t rec = idx / per_rec;
        int fld = (int)(idx - rec * per_rec);
        while (idx < want && next_token(k, b, e)) {
            if (!field(rec, fld, b, e)) {
                ok[t] = 0;
                return;
            }
            idx++;
            if (++fld == per_rec) {
                fld = 0;
                rec++;
            }
        }
        last[t] = k.p;
    };
    for (int t = 1; t < nthreads; t++) pool.emplace_back(parse_chunk, t);
    parse_chunk(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < nthreads; t++)
        if (!ok[t]) return false;
    // resume after the chunk that consumed the last wanted token
    int tl = 0;
    while (tl + 1 < nthreads && first[tl + 1] < want) tl++;
    c.p = last[tl];
    return true;
}

bool expect(Cursor& c, const char* word)
{
    const char *b, *e;
    if (!next_token(c, b, e)) return false;
    return (size_t)(e - b) == strlen(word) && memcmp(b, word, e - b) == 0;
}
bool next_int(Cursor& c, int64_t& v)
{
    const char *b, *e;
    return next_token(c, b, e) && to_i64(b, e, v);
}

// ModelFileT::GetElementSet -> nArray2DT::ReadNumbered (nArray2DT.h:1071): each record is "<number> n1 .. nen" and the row goes
// to index number - 1, whatever the order of the records (side sets address elements by that number).  A record can straddle a
// thread chunk, so rows are staged per record and scattered afterwards; numbers outside 1..nel and repeated numbers are errors.
bool parse_element_set(Cursor c, tb2_geom::Block& blk, int64_t nn, Cursor* after)
{
    int64_t nel, nen;
    if (!next_int(c, nel) || !next_int(c, nen) || nel != blk.nel || nen != blk.nen) return false;
    blk.conn.assign((size_t)nel * nen, 0);
    std::vector<int32_t> stage((size_t)nel * nen);
    std::vector<int64_t> row((size_t)nel, -1);
    int32_t* st = stage.data();
    int64_t* rowp = row.data();
    const int per = (int)nen + 1;
    bool okr = parse_records(c, nel, per, [=](int64_t rec, int fld, const char* b, const char* e) {
        int64_t v;
        if (!to_i64(b, e, v)) return false;
        if (fld == 0) {
            if (v < 1 || v > nel) return false;
            rowp[rec] = v - 1;
            return true;
        }
        if (v < 1 || v > nn) return false;
        st[rec * (per - 1) + fld - 1] = (int32_t)(v - 1);
        return true;
    });
    if (!okr) return false;
    std::vector<char> seen((size_t)nel, 0);
    int32_t* conn = blk.conn.data();
    for (int64_t r = 0; r < nel; r++) {
        const int64_t to = rowp[r];
        if (to < 0 || seen[to]) return false;
        seen[to] = 1;
        memcpy(conn + to * nen, st + r * nen, (size_t)nen * sizeof(int32_t));
    }
    if (after) *after = c;
    return true;
}

// ModelFileT::GetCoordinates -> dArray2DT::ReadNumbered: "<number> x y z", placed by number; repeated numbers are errors
bool parse_nodes(Cursor c, tb2_geom& g, Cursor* after)
{
    int64_t nn, nsd;
    if (!next_int(c, nn) || !next_int(c, nsd) || nn != g.nn || nsd != g.nsd) return false;
    g.X.assign((size_t)nn * 3, 0.0);
    double* X = g.X.data();
    const int per = (int)nsd + 1;
    std::vector<int64_t> row((size_t)nn, -1); // node id of each record: ids may come in any order
    int64_t* rowp = row.data();
    // two-step: ids first into row[], coordinates straight to their node -- a record's id precedes its coordinates in the same
    // chunk except when a chunk boundary splits the record, so coordinates are staged per record and scattered afterwards
    std::vector<double> stage((size_t)nn * nsd);
    double* st = stage.data();
    bool okr = parse_records(c, nn, per, [=](int64_t rec, int fld, const char* b, const char* e) {
        if (fld == 0) {
            int64_t v;
            if (!to_i64(b, e, v) || v < 1 || v > nn) return false;
            rowp[rec] = v - 1;
            return true;
        }
        double x;
        if (!to_f64(b, e, x)) return false;
        st[rec * (per - 1) + fld - 1] = x;
        return true;
    });
    if (!okr) return false;
    std::vector<char> seen((size_t)nn, 0);
    for (int64_t r = 0; r < nn; r++) {
        if (rowp[r] < 0 || seen[rowp[r]]) return false;
        seen[rowp[r]] = 1;
        for (int j = 0; j < (int)nsd; j++) X[rowp[r] * 3 + j] = st[r * nsd + j];
    }
    if (after) *after = c;
    return true;
}

int fail(const char* what, const std::string& path)
{
    set_error("%s: %s", path.c_str(), what);
    return TB2_ERR_ARG;
}

} // namespace

extern "C" {

int tb2_geom_open(const char* path, tb2_geom** out)
{
    TB2_ARG(path && out);
    *out = nullptr;
    std::vector<char> buf;
    const std::string file(path);
    PhaseTimer timer;
    if (!read_file(file, buf)) return fail("cannot read file", file);
    timer.mark("read file");
    const std::string dir = file.find_last_of('/') == std::string::npos ? std::string(".") : file.substr(0, file.find_last_of('/'));
    std::unique_ptr<tb2_geom> g(new tb2_geom);
    Cursor c{buf.data(), buf.data() + buf.size()};
    const char *b, *e;
    int64_t v;
    // *version <v>  *title <one free-text line>
    if (!expect(c, "*version") || !next_token(c, b, e)) return fail("expected *version", file);
    if (!expect(c, "*title")) return fail("expected *title", file);
    while (c.p < c.end && *c.p != '\n') c.p++; // rest of the keyword line
    if (c.p < c.end) c.p++;
    while (c.p < c.end && *c.p != '\n') c.p++; // the title line itself
    if (!expect(c, "*dimensions")) return fail("expected *dimensions", file);
    int64_t nblocks, nns, nss;
    if (!next_int(c, g->nn) || !next_int(c, v) || !next_int(c, nblocks)) return fail("bad *dimensions", file);
    g->nsd = (int)v;
    if (g->nn < 0 || g->nsd < 1 || g->nsd > 3 || nblocks < 0) return fail("bad *dimensions", file);
    g->blocks.resize((size_t)nblocks);
    for (auto& blk : g->blocks) {
        int64_t id, nel, nen;
        if (!next_int(c, id) || !next_int(c, nel) || !next_int(c, nen) || nel < 0 || nen < 1) return fail("bad element set dimensions", file);
        blk.id = (int)id;
        blk.nel = nel;
        blk.nen = (int)nen;
    }
    if (!next_int(c, nns) || nns < 0) return fail("bad node set count", file);
    std::vector<int64_t> ns_len((size_t)nns);
    g->nodesets.resize((size_t)nns);
    for (int64_t k = 0; k < nns; k++) {
        int64_t id;
        if (!next_int(c, id) || !next_int(c, ns_len[k]) || ns_len[k] < 0) return fail("bad node set dimensions", file);
        g->nodesets[k].id = (int)id;
    }
    if (!next_int(c, nss) || nss < 0) return fail("bad side set count", file);
    std::vector<int64_t> ss_len((size_t)nss);
    g->sidesets.resize((size_t)nss);
    for (int64_t k = 0; k < nss; k++) {
        int64_t id, blk;
        if (!next_int(c, id) || !next_int(c, blk) || !next_int(c, ss_len[k]) || ss_len[k] < 0) return fail("bad side set dimensions", file);
        g->sidesets[k].id = (int)id;
        g->sidesets[k].block = (int)blk;
    }
    // a set's body is inline, or the name of a file (relative to the main file) that holds it -- as for element sets and nodes
    // (ModelFileT::GetNodeSet / GetSideSet follow the same external-file convention, tri3.geom, quad4.*.geom)
    auto set_body = [&](Cursor& cur, std::vector<char>& ext_buf, Cursor& body, std::string& where) -> bool {
        Cursor look = cur;
        const char *tb, *te;
        int64_t tmp;
        if (!next_token(look, tb, te)) return false;
        if (to_i64(tb, te, tmp)) { // inline: the body starts here and the main cursor moves along with it
            body = cur;
            where = file;
            return true;
        }
        cur = look;
        where = dir + "/" + std::string(tb, te);
        if (!read_file(where, ext_buf)) return false;
        body = Cursor{ext_buf.data(), ext_buf.data() + ext_buf.size()};
        return true;
    };
    if (!expect(c, "*nodesets")) return fail("expected *nodesets", file);
    for (int64_t k = 0; k < nns; k++) {
        if (!expect(c, "*set")) return fail("bad node set", file);
        std::vector<char> ext_buf;
        Cursor body{nullptr, nullptr};
        std::string where;
        if (!set_body(c, ext_buf, body, where)) return fail("cannot read node set", where.empty() ? file : where);
        const bool inline_body = where == file;
        if (!next_int(body, v) || v != ns_len[k]) return fail("bad node set", where);
        auto& nodes = g->nodesets[k].nodes;
        nodes.resize((size_t)v);
        for (auto& n : nodes) {
            int64_t id;
            if (!next_int(body, id) || id == 0 || id > g->nn) return fail("bad node set entry", where);
            n = id < 0 ? -1 : (int32_t)(id - 1); // a negative entry is the reference's "all model nodes" marker (beam.1.geom): kept as -1
        }
        if (inline_body) c = body;
    }
    if (!expect(c, "*sidesets")) return fail("expected *sidesets", file);
    for (int64_t k = 0; k < nss; k++) {
        if (!expect(c, "*set")) return fail("bad side set", file);
        std::vector<char> ext_buf;
        Cursor body{nullptr, nullptr};
        std::string where;
        if (!set_body(c, ext_buf, body, where)) return fail("cannot read side set", where.empty() ? file : where);
        const bool inline_body = where == file;
        if (!next_int(body, v) || v != ss_len[k]) return fail("bad side set", where);
        auto& sides = g->sidesets[k].sides;
        sides.resize((size_t)v * 2);
        for (int64_t s = 0; s < v; s++) {
            int64_t el, fc;
            if (!next_int(body, el) || !next_int(body, fc) || el < 1 || fc < 1) return fail("bad side set entry", where);
            sides[2 * s] = (int32_t)(el - 1);
            sides[2 * s + 1] = (int32_t)(fc - 1);
        }
        if (inline_body) c = body;
    }
    timer.mark("dimensions, node/side sets");
    if (!expect(c, "*elements")) return fail("expected *elements", file);
    for (auto& blk : g->blocks) {
        if (!expect(c, "*set")) return fail("expected *set in *elements", file);
        Cursor look = c;
        if (!next_token(look, b, e)) return fail("truncated *elements", file);
        if (to_i64(b, e, v)) { // inline
            if (!parse_element_set(c, blk, g->nn, &c)) return fail("bad element set", file);
        } else { // external file, relative to the main file
            c = look;
            const std::string ext = dir + "/" + std::string(b, e);
            std::vector<char> eb;
            if (!read_file(ext, eb)) return fail("cannot read element file", ext);
            if (!parse_element_set(Cursor{eb.data(), eb.data() + eb.size()}, blk, g->nn, nullptr)) return fail("bad element set", ext);
        }
    }
    timer.mark("element sets");
    if (!expect(c, "*nodes")) return fail("expected *nodes", file);
    {
        Cursor look = c;
        if (!next_token(look, b, e)) return fail("truncated *nodes", file);
        if (to_i64(b, e, v)) {
            if (!parse_nodes(c, *g, &c)) return fail("bad node coordinates", file);
        } else {
            const std::string ext = dir + "/" + std::string(b, e);
            std::vector<char> nb;
            if (!read_file(ext, nb)) return fail("cannot read node file", ext);
            if (!parse_nodes(Cursor{nb.data(), nb.data() + nb.size()}, *g, nullptr)) return fail("bad node coordinates", ext);
        }
    }
    timer.mark("node coordinates");
    *out = g.release();
    return TB2_OK;
}

int tb2_geom_close(tb2_geom* g)
{
    delete g;
    return TB2_OK;
}

int tb2_geom_sizes(const tb2_geom* g, int64_t* nn, int32_t* nsd, int32_t* nblocks, int32_t* nnodesets, int32_t* nsidesets)
{
    TB2_ARG(g);
    if (nn) *nn = g->nn;
    if (nsd) *nsd = g->nsd;
    if (nblocks) *nblocks = (int32_t)g->blocks.size();
    if (nnodesets) *nnodesets = (int32_t)g->nodesets.size();
    if (nsidesets) *nsidesets = (int32_t)g->sidesets.size();
    return TB2_OK;
}

int tb2_geom_coords(const tb2_geom* g, double* X)
{
    TB2_ARG(g && X);
    memcpy(X, g->X.data(), g->X.size() * sizeof(double));
    return TB2_OK;
}

int tb2_geom_block(const tb2_geom* g, int32_t k, int32_t* id, int64_t* nel, int32_t* nen, int32_t* conn)
{
    TB2_ARG(g && k >= 0 && k < (int32_t)g->blocks.size());
    const auto& blk = g->blocks[k];
    if (id) *id = blk.id;
    if (nel) *nel = blk.nel;
    if (nen) *nen = blk.nen;
    if (conn) memcpy(conn, blk.conn.data(), blk.conn.size() * sizeof(int32_t));
    return TB2_OK;
}

int tb2_geom_nodeset(const tb2_geom* g, int32_t k, int32_t* id, int64_t* n, int32_t* nodes)
{
    TB2_ARG(g && k >= 0 && k < (int32_t)g->nodesets.size());
    const auto& s = g->nodesets[k];
    if (id) *id = s.id;
    if (n) *n = (int64_t)s.nodes.size();
    if (nodes) memcpy(nodes, s.nodes.data(), s.nodes.size() * sizeof(int32_t));
    return TB2_OK;
}

int tb2_geom_sideset(const tb2_geom* g, int32_t k, int32_t* id, int32_t* block_id, int64_t* n, int32_t* sides)
{
    TB2_ARG(g && k >= 0 && k < (int32_t)g->sidesets.size());
    const auto& s = g->sidesets[k];
    if (id) *id = s.id;
    if (block_id) *block_id = s.block;
    if (n) *n = (int64_t)s.sides.size() / 2;
    if (sides) memcpy(sides, s.sides.data(), s.sides.size() * sizeof(int32_t));
    return TB2_OK;
}

} // extern "C"
