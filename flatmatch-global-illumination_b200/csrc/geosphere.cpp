// Geodesic half-sphere direction sets (z > 0) for the ambient-occlusion pass.
//
// The reference ships them as generated constant tables (geoSphere.c, produced by geoSphere.py) and
// only its AO back-end uses one of them (geoSphere4, 481 directions, photonmap.c:450-453).  This
// regenerates the same point SET with the same construction - an octahedron's four upper faces,
// every triangle split in four `iterations` times, mid-points pushed onto the unit sphere
// (geoSphere.py:29-58), all arithmetic in double - instead of carrying the table.  The order of the
// directions differs from the reference's table (which reflects a Python dict's hash order); the AO
// sum is order-independent up to float rounding.
#include <cmath>
#include <cstring>
#include <vector>

#include "scene_tables.h"

namespace fmgi {

namespace {

struct D3 { double x, y, z; };
inline D3 mid_on_sphere(D3 a, D3 b)
{
    // geoSphere.py:44-46: normalized(div_vec3(add(a, b), 2.0))
    D3 m = {(a.x + b.x) / 2.0, (a.y + b.y) / 2.0, (a.z + b.z) / 2.0};
    const double len = sqrt(m.x * m.x + m.y * m.y + m.z * m.z);
    return {m.x / len, m.y / len, m.z / len};
}

void insert_unique(std::vector<D3> &set, D3 v)
{
    for (const D3 &u : set)
        if (u.x == v.x && u.y == v.y && u.z == v.z) return;     // exact match, as a Python dict key would
    set.push_back(v);
}

void subdivide(std::vector<D3> &set, D3 v1, D3 v2, D3 v3, int iterations)
{
    if (iterations <= 0) return;
    const D3 v12 = mid_on_sphere(v1, v2), v23 = mid_on_sphere(v2, v3), v31 = mid_on_sphere(v3, v1);
    if (iterations == 1) {
        const D3 all[6] = {v1, v2, v3, v12, v23, v31};
        for (const D3 &v : all) insert_unique(set, v);
        return;
    }
    subdivide(set, v1, v12, v31, iterations - 1);
    subdivide(set, v2, v12, v23, iterations - 1);
    subdivide(set, v3, v23, v31, iterations - 1);
    subdivide(set, v12, v23, v31, iterations - 1);
}

}  // namespace

// iterations 2, 3, 4, 5 give the reference's geoSphere2..5 (19, 113, 481, 1985 directions).
std::vector<float> geosphere_directions(int iterations)
{
    const double pi = 3.141592653589793;
    const D3 v1 = {0, 0, 1};
    D3 eq[4];
    for (int i = 0; i < 4; i++) {                               // geoSphere.py:61-64
        const double a = (90.0 * (i + 1)) / 180 * pi;
        eq[i] = {sin(a), cos(a), 0};
    }
    std::vector<D3> set;
    subdivide(set, v1, eq[0], eq[1], iterations);
    subdivide(set, v1, eq[1], eq[2], iterations);
    subdivide(set, v1, eq[2], eq[3], iterations);
    subdivide(set, v1, eq[3], eq[0], iterations);
    std::vector<float> out;
    for (const D3 &v : set)
        if (v.z != 0.0) {                                       // geoSphere.py:75: drop the equator
            out.push_back((float)v.x); out.push_back((float)v.y); out.push_back((float)v.z);
        }
    return out;
}

}  // namespace fmgi
