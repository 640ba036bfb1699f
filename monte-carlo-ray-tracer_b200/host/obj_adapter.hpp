// Drop-in bodies for Scene::parseOBJ and Scene::generateVertexNormals (source/scene/scene.cpp:238-355)
// on top of the parallel loader (obj_loader.hpp): same signatures, same containers, same results and
// the same failure behaviour (a missing file prints "<path> not found." and leaves the lists empty;
// a negative index throws std::runtime_error). A maintainer replaces the two function bodies with
//     mcrt_host::parseOBJ(path, vertices, normals, triangles_v, triangles_vt, triangles_vn);
//     mcrt_host::generateVertexNormals(normals, vertices, triangles);
#pragma once

#include <filesystem>
#include <iostream>
#include <stdexcept>
#include <vector>

#include <glm/vec3.hpp>

#include "obj_loader.hpp"

namespace mcrt_host
{
    inline void parseOBJ(const std::filesystem::path& path, std::vector<glm::dvec3>& vertices, std::vector<glm::dvec3>& normals,
                         std::vector<std::vector<size_t>>& triangles_v, std::vector<std::vector<size_t>>& triangles_vt,
                         std::vector<std::vector<size_t>>& triangles_vn)
    {
        if (!std::filesystem::exists(path))
        {
            std::cout << std::endl << path.string() << " not found.\n";
            return;
        }
        ObjMesh mesh;
        if (!parseOBJ(path.string(), mesh)) throw std::runtime_error(mesh.error);
        auto points = [](const std::vector<double>& src, std::vector<glm::dvec3>& dst)
        {
            dst.reserve(dst.size() + src.size() / 3);
            for (size_t i = 0; i + 2 < src.size(); i += 3) dst.emplace_back(src[i], src[i + 1], src[i + 2]);
        };
        auto corners = [](const std::vector<uint64_t>& src, std::vector<std::vector<size_t>>& dst)
        {
            dst.reserve(dst.size() + src.size() / 3);
            for (size_t i = 0; i + 2 < src.size(); i += 3) dst.push_back({ (size_t)src[i], (size_t)src[i + 1], (size_t)src[i + 2] });
        };
        points(mesh.vertices, vertices); points(mesh.normals, normals);
        corners(mesh.tri_v, triangles_v); corners(mesh.tri_vt, triangles_vt); corners(mesh.tri_vn, triangles_vn);
    }

    inline void generateVertexNormals(std::vector<glm::dvec3>& normals, const std::vector<glm::dvec3>& vertices,
                                      const std::vector<std::vector<size_t>>& triangles)
    {
        std::vector<double> v(3 * vertices.size()), n;
        for (size_t i = 0; i < vertices.size(); i++) { v[3 * i] = vertices[i].x; v[3 * i + 1] = vertices[i].y; v[3 * i + 2] = vertices[i].z; }
        std::vector<uint64_t> t;
        t.reserve(3 * triangles.size());
        for (const auto& tri : triangles) { t.push_back(tri.at(0)); t.push_back(tri.at(1)); t.push_back(tri.at(2)); }
        if (!generateVertexNormals(v, t, n)) throw std::out_of_range("generateVertexNormals: vertex index out of range");   // vector::at
        normals.resize(vertices.size());
        for (size_t i = 0; i < vertices.size(); i++) normals[i] = glm::dvec3(n[3 * i], n[3 * i + 1], n[3 * i + 2]);
    }
}
