// See obj_loader.hpp.
#include "obj_loader.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mcrt_host
{
namespace
{
    inline bool isSpace(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }   // std::isspace, "C" locale

    // operator>>(double): skip white space, take the characters num_get accumulates for a floating
    // literal ([+-] digits [. digits] [e|E [+-] digits]) and convert them. Failure (no digits) gives 0
    // and, as in a failed stream, stops the remaining extractions of the line (they keep 0 here; the
    // reference leaves them uninitialised).
    struct Cursor
    {
        const char* p; const char* end; bool failed = false;
        void skipSpace() { while (p < end && isSpace(*p)) p++; }
        double number()
        {
            if (failed) return 0.0;
            skipSpace();
            const char* s = p;
            const char* q = p;
            if (q < end && (*q == '+' || *q == '-')) q++;
            const char* digits = q;
            while (q < end && *q >= '0' && *q <= '9') q++;
            if (q < end && *q == '.') { q++; while (q < end && *q >= '0' && *q <= '9') q++; }
            const bool has_digits = (q - digits) > 1 || (q - digits == 1 && *digits != '.');
            if (!has_digits) { failed = true; return 0.0; }
            if (q < end && (*q == 'e' || *q == 'E'))
            {
                const char* r = q + 1;
                if (r < end && (*r == '+' || *r == '-')) r++;
                if (r < end && *r >= '0' && *r <= '9') { while (r < end && *r >= '0' && *r <= '9') r++; q = r; }
            }
            char buf[128];
            const size_t len = std::min<size_t>((size_t)(q - s), sizeof(buf) - 1);
            std::memcpy(buf, s, len); buf[len] = 0;
            p = q;
            return std::strtod(buf, nullptr);
        }
        // next white-space delimited token (ss >> std::string); empty at end of line
        void token(const char*& b, const char*& e)
        {
            skipSpace();
            b = p;
            while (p < end && !isSpace(*p)) p++;
            e = p;
        }
    };

    struct Chunk
    {
        std::vector<double> vertices, normals;
        std::vector<uint64_t> tri_v, tri_vt, tri_vn;
        bool negative = false;
    };

    // one corner "v", "v/vt", "v/vt/vn", "v//vn": std::getline(ss_f, idx, '/') pieces (scene.cpp:287-314)
    inline bool corner(const char* b, const char* e, uint64_t idx[4], int& count)
    {
        count = 0;
        const char* s = b;
        while (true)
        {
            const char* t = s;
            while (t < e && *t != '/') t++;
            if (s < e && *s == '-') return false;          // ss_f.peek() == '-'
            uint64_t v = 0;
            bool number = t > s;
            for (const char* c = s; c < t; c++) { if (*c < '0' || *c > '9') { number = false; break; } }
            if (number) for (const char* c = s; c < t; c++) v = v * 10 + (uint64_t)(*c - '0');   // std::stoull
            if (count < 4) idx[count] = number ? v : 0;
            count++;
            if (t >= e) break;
            s = t + 1;
            if (s >= e)
            {
                // "12/": after the delimiter the stream is still good(); peek hits EOF and the failed
                // getline leaves an empty piece
                if (count < 4) idx[count] = 0;
                count++;
                break;
            }
        }
        return true;
    }

    void parseRange(const char* begin, const char* end, Chunk& out)
    {
        const char* line = begin;
        while (line < end)
        {
            const char* eol = static_cast<const char*>(std::memchr(line, '\n', (size_t)(end - line)));
            if (!eol) eol = end;
            Cursor c{line, eol};
            const char *tb, *te;
            c.token(tb, te);
            const size_t tl = (size_t)(te - tb);
            if (tl == 1 && tb[0] == 'v')
            {
                const double x = c.number(), y = c.number(), z = c.number();
                out.vertices.push_back(x); out.vertices.push_back(y); out.vertices.push_back(z);
            }
            else if (tl == 2 && tb[0] == 'v' && tb[1] == 'n')
            {
                const double x = c.number(), y = c.number(), z = c.number();
                out.normals.push_back(x); out.normals.push_back(y); out.normals.push_back(z);
            }
            else if (tl == 1 && tb[0] == 'f')
            {
                uint64_t tv[3], tvt[3], tvn[3];
                int nv = 0, nvt = 0, nvn = 0;
                for (int i = 0; i < 3; i++)
                {
                    const char *b, *e;
                    c.token(b, e);
                    uint64_t idx[4]; int count;
                    if (!corner(b, e, idx, count)) { out.negative = true; return; }
                    if (count == 1) { tv[nv++] = idx[0] - 1; }
                    else if (count == 2) { tv[nv++] = idx[0] - 1; tvt[nvt++] = idx[1] - 1; }
                    else if (count == 3) { tv[nv++] = idx[0] - 1; if (idx[1]) tvt[nvt++] = idx[1] - 1; tvn[nvn++] = idx[2] - 1; }
                }
                if (nv == 3) out.tri_v.insert(out.tri_v.end(), tv, tv + 3);
                if (nvt == 3) out.tri_vt.insert(out.tri_vt.end(), tvt, tvt + 3);
                if (nvn == 3) out.tri_vn.insert(out.tri_vn.end(), tvn, tvn + 3);
            }
            line = eol + 1;
        }
    }

    int threadCount(int threads)
    {
        if (threads > 0) return threads;
        const unsigned hc = std::thread::hardware_concurrency();
        return hc ? (int)hc : 4;
    }

    template <class T> void appendAll(std::vector<T>& dst, const std::vector<Chunk>& chunks, std::vector<T> Chunk::*member)
    {
        size_t total = 0;
        for (const auto& c : chunks) total += (c.*member).size();
        dst.reserve(total);
        for (const auto& c : chunks) dst.insert(dst.end(), (c.*member).begin(), (c.*member).end());
    }
}

bool parseOBJ(const std::string& path, ObjMesh& out, int threads)
{
    out = ObjMesh();
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { out.error = path + " not found."; return false; }
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); out.error = path + " not found."; return false; }
    const size_t size = (size_t)st.st_size;
    if (size == 0) { ::close(fd); return true; }
    void* map = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) { out.error = "mmap failed: " + path; return false; }
    const char* data = static_cast<const char*>(map);

    int n = threadCount(threads);
    if (size < (1u << 16)) n = 1;
    std::vector<size_t> cut(n + 1);
    cut[0] = 0; cut[n] = size;
    for (int i = 1; i < n; i++)
    {
        size_t p = size / n * i;
        if (p < cut[i - 1]) p = cut[i - 1];
        const char* nl = static_cast<const char*>(std::memchr(data + p, '\n', size - p));
        cut[i] = nl ? (size_t)(nl - data) + 1 : size;
    }
    std::vector<Chunk> chunks(n);
    std::vector<std::thread> pool;
    for (int i = 1; i < n; i++) pool.emplace_back([&, i]() { parseRange(data + cut[i], data + cut[i + 1], chunks[i]); });
    parseRange(data + cut[0], data + cut[1], chunks[0]);
    for (auto& t : pool) t.join();
    ::munmap(map, size);

    for (const auto& c : chunks)
        if (c.negative) { out = ObjMesh(); out.error = "OBJ files with negative offsets are not supported."; return false; }
    appendAll(out.vertices, chunks, &Chunk::vertices);
    appendAll(out.normals, chunks, &Chunk::normals);
    appendAll(out.tri_v, chunks, &Chunk::tri_v);
    appendAll(out.tri_vt, chunks, &Chunk::tri_vt);
    appendAll(out.tri_vn, chunks, &Chunk::tri_vn);
    return true;
}

namespace
{
    struct D3 { double x, y, z; };
    inline D3 sub(const D3& a, const D3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
    inline D3 mul(const D3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
    inline double dot(const D3& a, const D3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                  // glm::dot
    inline D3 cross(const D3& a, const D3& b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }   // glm::cross
    inline D3 normalize(const D3& v) { return mul(v, 1.0 / std::sqrt(dot(v, v))); }                            // v * inversesqrt(dot(v, v))
    inline double angleBetween(const D3& a, const D3& b) { return std::acos(dot(normalize(a), normalize(b))); }
}

bool generateVertexNormals(const std::vector<double>& vertices, const std::vector<uint64_t>& tri_v, std::vector<double>& normals, int threads)
{
    const size_t nv = vertices.size() / 3, nt = tri_v.size() / 3;
    for (uint64_t i : tri_v) if (i >= nv) return false;
    if (tri_v.size() > 0xFFFFFFFFull) return false;   // corner indices are kept in 32 bits
    int n = threadCount(threads);
    if (nv < 4096) n = 1;
    auto parallel = [&](size_t count, auto&& body)
    {
        std::vector<std::thread> pool;
        for (int i = 1; i < n; i++) pool.emplace_back([&, i] { body(count * i / n, count * (i + 1) / n); });
        body(0, count / n);
        for (auto& t : pool) t.join();
    };
    // Phase 1, per triangle (once, not once per corner): the three angle-weighted area normals the reference adds to the
    // triangle's vertices (scene.cpp:330-347), in its expressions.
    const D3* V = reinterpret_cast<const D3*>(vertices.data());
    std::vector<D3> contribution(tri_v.size());
    parallel(nt, [&](size_t t0, size_t t1)
    {
        for (size_t t = t0; t < t1; t++)
        {
            const D3 a = V[tri_v[3 * t]], b = V[tri_v[3 * t + 1]], c = V[tri_v[3 * t + 2]];
            // Surface::Triangle(v0, v1, v2): E1, E2, normal_ = normalize(cross(E1, E2)), area_ = |cross| / 2
            const D3 e1 = sub(b, a), e2 = sub(c, a), cr = cross(e1, e2);
            const D3 weighted = mul(normalize(cr), std::sqrt(dot(cr, cr)) / 2.0);
            contribution[3 * t] = mul(weighted, angleBetween(sub(a, b), sub(a, c)));
            contribution[3 * t + 1] = mul(weighted, angleBetween(sub(b, a), sub(b, c)));
            contribution[3 * t + 2] = mul(weighted, angleBetween(sub(c, a), sub(c, b)));
        }
    });
    // incidence lists in triangle order (counting sort by vertex): corner c of triangle t -> slot; the reference accumulates
    // triangle by triangle, so every vertex sums its corners in ascending triangle order
    std::vector<uint32_t> start(nv + 1, 0);
    for (uint64_t i : tri_v) start[i + 1]++;
    for (size_t v = 0; v < nv; v++) start[v + 1] += start[v];
    std::vector<uint32_t> incident(tri_v.size());
    {
        std::vector<uint32_t> cursor(start.begin(), start.end() - 1);
        for (size_t k = 0; k < tri_v.size(); k++) incident[cursor[tri_v[k]]++] = (uint32_t)k;   // k = 3 * triangle + corner, ascending
    }
    normals.assign(3 * nv, 0.0);
    // Phase 2, per vertex: sum and normalise
    parallel(nv, [&](size_t v0, size_t v1)
    {
        for (size_t v = v0; v < v1; v++)
        {
            D3 sum{0.0, 0.0, 0.0};
            for (uint32_t s = start[v]; s < start[v + 1]; s++)
            {
                const D3& add = contribution[incident[s]];
                sum.x += add.x; sum.y += add.y; sum.z += add.z;
            }
            const D3 nrm = normalize(sum);
            normals[3 * v] = nrm.x; normals[3 * v + 1] = nrm.y; normals[3 * v + 2] = nrm.z;
        }
    });
    return true;
}
}
