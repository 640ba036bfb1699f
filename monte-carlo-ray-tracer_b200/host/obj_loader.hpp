// Scene ingest (widened scope): the reference's OBJ reader and vertex-normal
// generator (source/scene/scene.cpp:238-355) as a parallel, mmap-based loader with the same results.
//   Scene::parseOBJ             -> parseOBJ:  vertices, normals, first three corners of every face
//   Scene::generateVertexNormals -> generateVertexNormals: angle- and area-weighted normals
// The reference reads the file line by line through iostreams on one thread (1.5 s for the
// 457 k-triangle spaceship); this loader splits the mapped file at line boundaries, parses the
// chunks on all cores and concatenates them in file order, so the arrays are identical. Numbers go
// through strtod on exactly the characters num_get would consume (both are correctly rounded).
// No dependency on the reference's headers or on CUDA.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace mcrt_host
{
    struct ObjMesh
    {
        std::vector<double> vertices;        // "v" lines, xyz
        std::vector<double> normals;         // "vn" lines, xyz
        // first three corners of every "f" line, indices already zero-based (idx - 1 in size_t
        // arithmetic, so a missing index wraps to SIZE_MAX as in the reference). As in the reference
        // the three lists are independent: a face contributes to tri_vt / tri_vn only when all three
        // of its corners carry that index.
        std::vector<uint64_t> tri_v, tri_vt, tri_vn;
        std::string error;                   // "OBJ files with negative offsets are not supported." (scene.cpp:295)
    };

    // false: file missing (the reference prints "<path> not found." and returns empty lists) or
    // negative index (the reference throws). threads <= 0: hardware concurrency.
    bool parseOBJ(const std::string& path, ObjMesh& out, int threads = 0);

    // normals[vertex] = normalize(sum over incident triangles, in triangle order, of
    // face_normal * area * corner_angle) (scene.cpp:326-355). Parallel over vertices; every vertex
    // adds its contributions in the reference's order, so the sums are bit-identical.
    // Returns false if a triangle references a vertex out of range (the reference's .at() throws).
    bool generateVertexNormals(const std::vector<double>& vertices, const std::vector<uint64_t>& tri_v,
                               std::vector<double>& normals, int threads = 0);
}
