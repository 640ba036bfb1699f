// See exporter.hpp. Compile with -fno-access-control against /root/reference/{source,lib/*}.
#include "exporter.hpp"

#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "scene/scene.hpp"
#include "bvh/bvh.hpp"
#include "surface/surface.hpp"
#include "material/material.hpp"
#include "material/fresnel.hpp"
#include "camera/camera.hpp"
#include "camera/filter.hpp"
#include "integrator/photon-mapper/photon-mapper.hpp"

namespace mcrt_host
{
    static void push3(std::vector<double>& v, const glm::dvec3& a)
    {
        v.push_back(a.x); v.push_back(a.y); v.push_back(a.z);
    }

    static void pushBounds(std::vector<double>& v, const BoundingBox& bb)
    {
        push3(v, bb.min); push3(v, bb.max);
    }

    static mcrt_material flattenMaterial(const Material& m)
    {
        mcrt_material o;
        std::memset(&o, 0, sizeof(o));
        for (int i = 0; i < 3; i++)
        {
            o.reflectance[i] = m.reflectance[i];
            o.specular_reflectance[i] = m.specular_reflectance[i];
            o.transmittance[i] = m.transmittance[i];
            o.emittance[i] = m.emittance[i];
        }
        o.roughness = m.roughness;
        o.specular_roughness = m.specular_roughness;
        o.ior = m.ior;
        o.transparency = m.transparency;
        if (m.complex_ior)
        {
            o.has_complex_ior = 1;
            for (int i = 0; i < 3; i++)
            {
                o.complex_ior_real[i] = m.complex_ior->real[i];
                o.complex_ior_imag[i] = m.complex_ior->imaginary[i];
            }
        }
        o.A = m.A; o.B = m.B;
        o.a[0] = m.a.x; o.a[1] = m.a.y;
        o.perfect_mirror = m.perfect_mirror;
        o.rough = m.rough;
        o.rough_specular = m.rough_specular;
        o.opaque = m.opaque;
        o.emissive = m.emissive;
        o.dirac_delta = m.dirac_delta;
        return o;
    }

    void flattenScene(const Scene& scene, FlatScene& out)
    {
        out = FlatScene();
        out.scene_ior = scene.ior;

        const std::vector<std::shared_ptr<Surface::Base>>* ordered = &scene.surfaces;
        if (scene.bvh)
        {
            const BVH& bvh = *scene.bvh;
            ordered = &bvh.ordered_surfaces;
            for (const auto& node : bvh.linear_tree)
            {
                pushBounds(out.node_bounds, node.BB);
                out.node_first_prim.push_back(node.start_surface);
                out.node_prim_count.push_back(node.num_surfaces);
                out.node_next_sibling.push_back(node.next_sibling);
            }
        }

        std::unordered_map<const Material*, uint32_t> material_index;
        auto materialOf = [&](const std::shared_ptr<Material>& m) -> uint32_t
        {
            auto it = material_index.find(m.get());
            if (it != material_index.end()) return it->second;
            uint32_t idx = (uint32_t)out.materials.size();
            out.materials.push_back(flattenMaterial(*m));
            material_index.emplace(m.get(), idx);
            return idx;
        };

        std::unordered_map<const Surface::Base*, uint32_t> original_index;
        for (size_t i = 0; i < scene.surfaces.size(); i++) original_index.emplace(scene.surfaces[i].get(), (uint32_t)i);
        {
            const BoundingBox bb = scene.BB();
            for (int c = 0; c < 3; c++) { out.scene_bounds[c] = bb.min[c]; out.scene_bounds[3 + c] = bb.max[c]; }
        }

        for (const auto& sp : *ordered)
        {
            const Surface::Base* s = sp.get();
            if (!s) throw std::runtime_error("exporter: null surface in ordered list");
            out.prim_original.push_back(original_index.at(s));
            uint32_t prim = (uint32_t)out.prim_type.size();
            out.prim_of_surface.emplace(s, prim);
            out.prim_material.push_back(materialOf(s->material));
            out.prim_area.push_back(s->area_);

            if (auto* t = dynamic_cast<const Surface::Triangle*>(s))
            {
                out.prim_type.push_back(MCRT_PRIM_TRIANGLE);
                out.prim_index.push_back((uint32_t)out.tri_vn_index.size());
                push3(out.tri_v0, t->v0); push3(out.tri_v1, t->v1); push3(out.tri_v2, t->v2);
                push3(out.tri_e1, t->E1); push3(out.tri_e2, t->E2); push3(out.tri_normal, t->normal_);
                if (t->N)
                {
                    out.tri_vn_index.push_back((int32_t)(out.vertex_normals.size() / 9));
                    for (int c = 0; c < 3; c++) push3(out.vertex_normals, (*t->N)[c]);
                }
                else
                {
                    out.tri_vn_index.push_back(-1);
                }
            }
            else if (auto* sph = dynamic_cast<const Surface::Sphere*>(s))
            {
                out.prim_type.push_back(MCRT_PRIM_SPHERE);
                out.prim_index.push_back((uint32_t)(out.sphere_origin_radius.size() / 4));
                push3(out.sphere_origin_radius, sph->origin);
                out.sphere_origin_radius.push_back(sph->radius);
            }
            else if (auto* q = dynamic_cast<const Surface::Quadric*>(s))
            {
                out.prim_type.push_back(MCRT_PRIM_QUADRIC);
                out.prim_index.push_back((uint32_t)(out.quadric_bounds.size() / 6));
                for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) out.quadric_Q.push_back(q->Q[c][r]);
                for (int c = 0; c < 4; c++) for (int r = 0; r < 3; r++) out.quadric_G.push_back(q->G[c][r]);
                pushBounds(out.quadric_bounds, q->BB_);
            }
            else
            {
                throw std::runtime_error("exporter: unknown surface type");
            }
        }

        for (size_t i = 0; i < scene.emissives.size(); i++)
        {
            auto it = out.prim_of_surface.find(scene.emissives[i].get());
            if (it == out.prim_of_surface.end()) throw std::runtime_error("exporter: emissive not in surfaces");
            out.light_prim.push_back(it->second);
            out.light_cdf.push_back(scene.cumulative_emissives_importance[i]);
        }
    }

    mcrt_scene_desc FlatScene::desc() const
    {
        mcrt_scene_desc d;
        std::memset(&d, 0, sizeof(d));
        d.abi_version = MCRT_ABI_VERSION;
        d.n_nodes = (uint32_t)node_first_prim.size();
        d.node_bounds = node_bounds.data();
        d.node_first_prim = node_first_prim.data();
        d.node_prim_count = node_prim_count.data();
        d.node_next_sibling = node_next_sibling.data();
        d.n_prims = (uint32_t)prim_type.size();
        d.prim_type = prim_type.data();
        d.prim_index = prim_index.data();
        d.prim_material = prim_material.data();
        d.prim_area = prim_area.data();
        d.n_tris = (uint32_t)tri_vn_index.size();
        d.tri_v0 = tri_v0.data(); d.tri_v1 = tri_v1.data(); d.tri_v2 = tri_v2.data();
        d.tri_e1 = tri_e1.data(); d.tri_e2 = tri_e2.data(); d.tri_normal = tri_normal.data();
        d.tri_vn_index = tri_vn_index.data();
        d.n_vertex_normals = (uint32_t)(vertex_normals.size() / 9);
        d.vertex_normals = vertex_normals.data();
        d.n_spheres = (uint32_t)(sphere_origin_radius.size() / 4);
        d.sphere_origin_radius = sphere_origin_radius.data();
        d.n_quadrics = (uint32_t)(quadric_bounds.size() / 6);
        d.quadric_Q = quadric_Q.data(); d.quadric_G = quadric_G.data(); d.quadric_bounds = quadric_bounds.data();
        d.n_materials = (uint32_t)materials.size();
        d.materials = materials.data();
        d.n_lights = (uint32_t)light_prim.size();
        d.light_prim = light_prim.data();
        d.light_cdf = light_cdf.data();
        d.scene_ior = scene_ior;
        return d;
    }

    mcrt_camera flattenCamera(const Camera& c)
    {
        mcrt_camera o;
        std::memset(&o, 0, sizeof(o));
        for (int i = 0; i < 3; i++)
        {
            o.eye[i] = c.eye[i]; o.forward[i] = c.forward[i]; o.left[i] = c.left[i]; o.up[i] = c.up[i];
        }
        o.focal_length = c.focal_length;
        o.sensor_width = c.sensor_width;
        o.aperture_radius = c.aperture_radius;
        o.focus_distance = c.focus_distance;
        o.width = (uint32_t)c.image.width;
        o.height = (uint32_t)c.image.height;
        o.thin_lens = c.thin_lens;
        return o;
    }

    void primitiveBounds(const Scene& scene, std::vector<double>& prim_bounds, double scene_bounds[6])
    {
        prim_bounds.clear();
        prim_bounds.reserve(6 * scene.surfaces.size());
        for (const auto& s : scene.surfaces) pushBounds(prim_bounds, s->BB());
        const BoundingBox bb = scene.BB();
        for (int c = 0; c < 3; c++) { scene_bounds[c] = bb.min[c]; scene_bounds[3 + c] = bb.max[c]; }
    }

    void applyBvh(FlatScene& f, const mcrt_bvh_desc& bvh)
    {
        const size_t n = f.prim_type.size();
        if (bvh.n_prims != n || !f.node_first_prim.empty()) throw std::runtime_error("applyBvh: scene/tree mismatch");
        std::vector<uint32_t> inverse(n);
        for (size_t i = 0; i < n; i++) inverse[bvh.prim_order[i]] = (uint32_t)i;
        auto permute = [&](auto& v)
        {
            auto old = v;
            for (size_t i = 0; i < n; i++) v[i] = old[bvh.prim_order[i]];
        };
        permute(f.prim_type); permute(f.prim_index); permute(f.prim_material); permute(f.prim_area); permute(f.prim_original);
        for (auto& l : f.light_prim) l = inverse[l];
        for (auto& kv : f.prim_of_surface) kv.second = inverse[kv.second];
        f.node_bounds.assign(bvh.node_bounds, bvh.node_bounds + 6 * (size_t)bvh.n_nodes);
        f.node_first_prim.assign(bvh.node_first_prim, bvh.node_first_prim + bvh.n_nodes);
        f.node_prim_count.assign(bvh.node_prim_count, bvh.node_prim_count + bvh.n_nodes);
        f.node_next_sibling.assign(bvh.node_next_sibling, bvh.node_next_sibling + bvh.n_nodes);
    }

    mcrt_film flattenFilm(const Camera& c)
    {
        // Film keeps its filter as a std::function built from a plain function (film.cpp:25-45);
        // the target's address names it.
        typedef double (*Fn)(double);
        const Fn* target = c.film.filter_function.target<Fn>();
        if (!target) throw std::runtime_error("flattenFilm: unknown film filter");
        mcrt_film f;
        if (*target == static_cast<Fn>(Filter::box)) f.filter = MCRT_FILM_BOX;
        else if (*target == static_cast<Fn>(Filter::MitchellNetravali<>)) f.filter = MCRT_FILM_MITCHELL_NETRAVALI;
        else if (*target == static_cast<Fn>(Filter::CatmullRom)) f.filter = MCRT_FILM_CATMULL_ROM;
        else if (*target == static_cast<Fn>(Filter::BSpline)) f.filter = MCRT_FILM_B_SPLINE;
        else if (*target == static_cast<Fn>(Filter::Hermite)) f.filter = MCRT_FILM_HERMITE;
        else if (*target == static_cast<Fn>(Filter::Gaussian)) f.filter = MCRT_FILM_GAUSSIAN;
        else if (*target == static_cast<Fn>(Filter::Lanczos)) f.filter = MCRT_FILM_LANCZOS;
        else throw std::runtime_error("flattenFilm: unknown film filter");
        f.cache_size = (uint32_t)c.film.filter_cache.size();
        f.radius = c.film.radius;
        return f;
    }

    void flattenPhotonMap(const PhotonMapper& pm, int which, FlatPhotonMap& out)
    {
        out = FlatPhotonMap();
        const LinearOctree<Photon>& map = which == 0 ? pm.caustic_map : pm.global_map;
        for (const auto& o : map.linear_tree)
        {
            pushBounds(out.octant_bounds, o.BB);
            out.octant_start.push_back(o.start_data);
            out.octant_count.push_back(o.contained_data);
            out.octant_next_sibling.push_back(o.next_sibling);
            out.octant_leaf.push_back(o.leaf);
        }
        out.photons.reserve(map.ordered_data.size() * 8);
        for (const auto& p : map.ordered_data)
        {
            for (int i = 0; i < 3; i++) out.photons.push_back(p.flux_[i]);
            for (int i = 0; i < 3; i++) out.photons.push_back(p.position_[i]);
            out.photons.push_back(p.phi);
            out.photons.push_back(p.theta);
        }
    }

    void photonMapParams(const PhotonMapper& pm, uint32_t& k_nearest, uint32_t& direct_visualization)
    {
        k_nearest = (uint32_t)pm.k_nearest_photons;
        direct_visualization = pm.direct_visualization;
    }

    mcrt_photon_map_desc FlatPhotonMap::desc() const
    {
        mcrt_photon_map_desc d;
        std::memset(&d, 0, sizeof(d));
        d.n_octants = (uint32_t)octant_leaf.size();
        d.octant_bounds = octant_bounds.data();
        d.octant_start = octant_start.data();
        d.octant_count = octant_count.data();
        d.octant_next_sibling = octant_next_sibling.data();
        d.octant_leaf = octant_leaf.data();
        d.n_photons = photons.size() / 8;
        d.photons = photons.data();
        return d;
    }

    // ------------------------------------------------------------------ pack file

    void PackWriter::addScalars(const std::string& name, const std::vector<double>& v)
    {
        owned.emplace_back((const uint8_t*)v.data(), (const uint8_t*)(v.data() + v.size()));
        entries.push_back({ name, 5u, 8u, v.size(), nullptr });
        entries.back().data = (const void*)(uintptr_t)(owned.size()); // resolved at write time
    }

    void PackWriter::addScalarsU32(const std::string& name, const std::vector<uint32_t>& v)
    {
        owned.emplace_back((const uint8_t*)v.data(), (const uint8_t*)(v.data() + v.size()));
        entries.push_back({ name, 1u, 4u, v.size(), nullptr });
        entries.back().data = (const void*)(uintptr_t)(owned.size());
    }

    bool PackWriter::write(const std::string& path) const
    {
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) return false;

        struct DiskEntry { char name[32]; uint32_t dtype, elem_size; uint64_t count, offset; };
        static_assert(sizeof(DiskEntry) == 56, "pack entry layout");

        std::vector<DiskEntry> table(entries.size());
        uint64_t offset = 8 + 4 + 4 + sizeof(DiskEntry) * entries.size();
        auto align64 = [](uint64_t x) { return (x + 63) & ~uint64_t(63); };
        offset = align64(offset);
        for (size_t i = 0; i < entries.size(); i++)
        {
            std::memset(&table[i], 0, sizeof(DiskEntry));
            std::strncpy(table[i].name, entries[i].name.c_str(), 31);
            table[i].dtype = entries[i].dtype;
            table[i].elem_size = entries[i].elem_size;
            table[i].count = entries[i].count;
            table[i].offset = offset;
            offset = align64(offset + entries[i].count * entries[i].elem_size);
        }

        const char magic[8] = { 'M','C','R','T','P','K','0','1' };
        uint32_t n = (uint32_t)entries.size(), zero = 0;
        std::fwrite(magic, 1, 8, f);
        std::fwrite(&n, 4, 1, f);
        std::fwrite(&zero, 4, 1, f);
        std::fwrite(table.data(), sizeof(DiskEntry), table.size(), f);

        for (size_t i = 0; i < entries.size(); i++)
        {
            std::fseek(f, (long)table[i].offset, SEEK_SET);
            const void* data = entries[i].data;
            uintptr_t tag = (uintptr_t)data;
            if (tag >= 1 && tag <= owned.size()) data = owned[tag - 1].data(); // owned scalar block
            size_t bytes = entries[i].count * entries[i].elem_size;
            if (bytes && std::fwrite(data, 1, bytes, f) != bytes) { std::fclose(f); return false; }
        }
        // pad the tail so that readers can map whole 64-byte blocks
        std::fseek(f, (long)offset - 1, SEEK_SET);
        char z = 0; std::fwrite(&z, 1, 1, f);
        std::fclose(f);
        return true;
    }

    void addSceneToPack(PackWriter& w, const FlatScene& s)
    {
        w.add("node_bounds", 5, s.node_bounds);
        w.add("node_first_prim", 1, s.node_first_prim);
        w.add("node_prim_count", 1, s.node_prim_count);
        w.add("node_next_sibling", 1, s.node_next_sibling);
        w.add("prim_type", 0, s.prim_type);
        w.add("prim_index", 1, s.prim_index);
        w.add("prim_material", 1, s.prim_material);
        w.add("prim_area", 5, s.prim_area);
        w.add("tri_v0", 5, s.tri_v0); w.add("tri_v1", 5, s.tri_v1); w.add("tri_v2", 5, s.tri_v2);
        w.add("tri_e1", 5, s.tri_e1); w.add("tri_e2", 5, s.tri_e2); w.add("tri_normal", 5, s.tri_normal);
        w.add("tri_vn_index", 2, s.tri_vn_index);
        w.add("vertex_normals", 5, s.vertex_normals);
        w.add("sphere_origin_radius", 5, s.sphere_origin_radius);
        w.add("quadric_Q", 5, s.quadric_Q); w.add("quadric_G", 5, s.quadric_G);
        w.add("quadric_bounds", 5, s.quadric_bounds);
        w.add("materials", 6, s.materials);
        w.add("light_prim", 1, s.light_prim);
        w.add("light_cdf", 5, s.light_cdf);
        w.addScalars("scene_ior", { s.scene_ior });
        w.add("prim_original", 1, s.prim_original);
        w.addScalars("scene_bounds", std::vector<double>(s.scene_bounds, s.scene_bounds + 6));
    }

    void addCameraToPack(PackWriter& w, const std::string& prefix, const mcrt_camera& c, uint32_t sqrtspp)
    {
        std::vector<double> d;
        for (int i = 0; i < 3; i++) d.push_back(c.eye[i]);
        for (int i = 0; i < 3; i++) d.push_back(c.forward[i]);
        for (int i = 0; i < 3; i++) d.push_back(c.left[i]);
        for (int i = 0; i < 3; i++) d.push_back(c.up[i]);
        d.push_back(c.focal_length); d.push_back(c.sensor_width);
        d.push_back(c.aperture_radius); d.push_back(c.focus_distance);
        w.addScalars(prefix + "_f64", d);
        w.addScalarsU32(prefix + "_u32", { c.width, c.height, c.thin_lens, sqrtspp });
    }

    void addFilmToPack(PackWriter& w, const std::string& prefix, const mcrt_film& f)
    {
        w.addScalarsU32(prefix + "_film_u32", { f.filter, f.cache_size });
        w.addScalars(prefix + "_film_f64", { f.radius });
    }

    void addPhotonMapToPack(PackWriter& w, const std::string& prefix, const FlatPhotonMap& m)
    {
        w.add(prefix + "_octant_bounds", 5, m.octant_bounds);
        w.add(prefix + "_octant_start", 3, m.octant_start);
        w.add(prefix + "_octant_count", 3, m.octant_count);
        w.add(prefix + "_octant_next", 1, m.octant_next_sibling);
        w.add(prefix + "_octant_leaf", 0, m.octant_leaf);
        w.add(prefix + "_photons", 4, m.photons);
    }
}
