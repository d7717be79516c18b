// ORACLE (test infrastructure only -- never linked into the product path).
#include "robot_desc.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace oracle {

static bool fail(std::string* err, const std::string& msg)
{
    if (err) {
        *err = msg;
    }
    return false;
}

bool LoadRobotDesc(const std::string& path, RobotDesc& d, std::string* err)
{
    std::ifstream f(path.c_str());
    if (!f) {
        return fail(err, "cannot open " + path);
    }
    d = RobotDesc();
    std::string line;
    int lineno = 0;
    while (std::getline(f, line)) {
        ++lineno;
        std::istringstream ss(line);
        std::string key;
        if (!(ss >> key) || key[0] == '#') {
            continue;
        }
        auto bad = [&]() {
            return fail(err, path + ":" + std::to_string(lineno) + ": malformed '" + key + "' record");
        };
        if (key == "robot") {
            if (!(ss >> d.name)) return bad();
        } else if (key == "root") {
            if (!(ss >> d.root)) return bad();
        } else if (key == "world_joint") {
            if (!(ss >> d.world_joint_name >> d.world_joint_type)) return bad();
        } else if (key == "joint") {
            JointDesc j;
            int hl, hs;
            if (!(ss >> j.name >> j.type >> j.parent >> j.child
                     >> j.xyz[0] >> j.xyz[1] >> j.xyz[2]
                     >> j.rpy[0] >> j.rpy[1] >> j.rpy[2]
                     >> j.axis[0] >> j.axis[1] >> j.axis[2]
                     >> hl >> j.lower >> j.upper >> hs >> j.soft_lower >> j.soft_upper))
            {
                return bad();
            }
            j.has_limits = hl != 0;
            j.has_safety = hs != 0;
            d.joints.push_back(j);
        } else if (key == "spheres_model") {
            SpheresModelConfig m;
            if (!(ss >> m.link_name)) return bad();
            d.spheres_models.push_back(m);
        } else if (key == "sphere") {
            SphereConfig s;
            if (d.spheres_models.empty()) return bad();
            if (!(ss >> s.name >> s.x >> s.y >> s.z >> s.radius >> s.priority)) return bad();
            d.spheres_models.back().spheres.push_back(s);
        } else if (key == "voxels_model") {
            VoxelsModelConfig v;
            if (!(ss >> v.link_name >> v.res >> v.center[0] >> v.center[1] >> v.center[2]
                     >> v.size[0] >> v.size[1] >> v.size[2]))
            {
                return bad();
            }
            d.voxels_models.push_back(v);
        } else if (key == "group") {
            GroupConfig g;
            if (!(ss >> g.name)) return bad();
            d.groups.push_back(g);
        } else if (key == "group_link") {
            std::string n;
            if (d.groups.empty() || !(ss >> n)) return bad();
            d.groups.back().links.push_back(n);
        } else if (key == "group_chain") {
            std::string b, t;
            if (d.groups.empty() || !(ss >> b >> t)) return bad();
            d.groups.back().chains.push_back(std::make_pair(b, t));
        } else if (key == "group_sub") {
            std::string n;
            if (d.groups.empty() || !(ss >> n)) return bad();
            d.groups.back().groups.push_back(n);
        } else if (key == "acm") {
            AcmEntryDesc e;
            int a;
            if (!(ss >> e.a >> e.b >> a)) return bad();
            e.allowed = a != 0;
            d.acm.push_back(e);
        } else {
            return bad();
        }
    }
    if (d.name.empty() || d.root.empty()) {
        return fail(err, path + ": missing robot/root record");
    }
    return true;
}

} // namespace oracle
