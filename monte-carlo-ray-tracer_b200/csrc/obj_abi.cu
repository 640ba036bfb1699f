// C ABI of the scene-ingest helpers: host code only, compiled into the same
// library so that one binding covers the whole boundary. See host/obj_loader.hpp.
#include <cstring>
#include <new>

#include "../../include/mcrt_abi.h"
#include "../host/obj_loader.cpp"

extern "C"
{

int mcrt_obj_load(const char* path, int threads, void** handle, mcrt_obj_mesh* out, char* err, size_t errlen)
{
    if (!path || !handle || !out) return MCRT_ERR_INVALID;
    *handle = nullptr;
    std::memset(out, 0, sizeof(*out));
    auto* m = new (std::nothrow) mcrt_host::ObjMesh();
    if (!m) return MCRT_ERR_INVALID;
    if (!mcrt_host::parseOBJ(path, *m, threads))
    {
        if (err && errlen) { std::strncpy(err, m->error.c_str(), errlen - 1); err[errlen - 1] = 0; }
        delete m;
        return MCRT_ERR_INVALID;
    }
    out->n_vertices = m->vertices.size() / 3; out->n_normals = m->normals.size() / 3;
    out->n_tri_v = m->tri_v.size() / 3; out->n_tri_vt = m->tri_vt.size() / 3; out->n_tri_vn = m->tri_vn.size() / 3;
    out->vertices = m->vertices.data(); out->normals = m->normals.data();
    out->tri_v = m->tri_v.data(); out->tri_vt = m->tri_vt.data(); out->tri_vn = m->tri_vn.data();
    *handle = m;
    return MCRT_OK;
}

void mcrt_obj_free(void* handle)
{
    delete static_cast<mcrt_host::ObjMesh*>(handle);
}

int mcrt_obj_vertex_normals(const double* vertices, uint64_t n_vertices, const uint64_t* tri_v, uint64_t n_triangles, int threads,
                            double* out_normals)
{
    if ((n_vertices && !vertices) || (n_triangles && !tri_v) || (n_vertices && !out_normals)) return MCRT_ERR_INVALID;
    std::vector<double> v(vertices, vertices + 3 * n_vertices), n;
    std::vector<uint64_t> t(tri_v, tri_v + 3 * n_triangles);
    if (!mcrt_host::generateVertexNormals(v, t, n, threads)) return MCRT_ERR_INVALID;
    if (n_vertices) std::memcpy(out_normals, n.data(), n.size() * sizeof(double));
    return MCRT_OK;
}

}
