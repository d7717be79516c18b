"""ctypes binding of the CPU oracle (oracle/liboracle.so, oracle/_ref/libref_bfs3d.so).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module; nothing under
smpl_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None
_REF = None

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


def build_oracle():
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "-j8"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, stdout=subprocess.DEVNULL)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


def _bp(a):
    return a.ctypes.data_as(c_uint8_p)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_scene_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, c_double_p, c_double_p,
                                          C.c_double, C.c_double]
        L.oracle_bfs_create.restype = C.c_void_p
        L.oracle_bfs_create.argtypes = [C.c_int, C.c_int, C.c_int]
        for name in ("oracle_time_states_valid", "oracle_time_edges_valid", "oracle_time_bfs_run", "oracle_plan", "oracle_plan_lazy"):
            getattr(L, name).restype = C.c_double
        L.oracle_time_bfs_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own BFS_3D (compiled from /root/reference), or None."""
    global _REF
    if _REF is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_bfs3d.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_bfs_create.restype = C.c_void_p
        R.ref_bfs_create.argtypes = [C.c_int, C.c_int, C.c_int]
        R.ref_bfs_time_run.restype = C.c_double
        R.ref_bfs_time_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _REF = R
    return _REF


_REF_DM = None


def ref_distmap_lib():
    """The reference's own EuclidDistanceMap (compiled from /root/reference), or None."""
    global _REF_DM
    if _REF_DM is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_distmap.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_distmap_create.restype = C.c_void_p
        R.ref_distmap_create.argtypes = [C.c_double] * 8
        R.ref_distmap_distance.restype = C.c_double
        R.ref_distmap_distance.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        R.ref_distmap_world_to_grid.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, c_int32_p]
        _REF_DM = R
    return _REF_DM


class RefDistanceMap:
    """sbpl::EuclidDistanceMap of the reference (with the fork's commented-out interior initialisation restored
    by the shim, oracle/ref_distmap_shim.cpp)."""

    def __init__(self, origin, size, res, max_dist):
        self.R = ref_distmap_lib()
        self.h = C.c_void_p(self.R.ref_distmap_create(*(float(v) for v in (*origin, *size, res, max_dist))))
        d = np.zeros(3, np.int32)
        self.R.ref_distmap_dims(self.h, _ip(d))
        self.dims = tuple(int(v) for v in d)

    def close(self):
        if self.h:
            self.R.ref_distmap_destroy(self.h)
            self.h = None

    def add_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self.R.ref_distmap_add_points(self.h, _dp(pts), len(pts))

    def remove_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self.R.ref_distmap_remove_points(self.h, _dp(pts), len(pts))

    def d2(self):
        out = np.zeros(self.dims[0] * self.dims[1] * self.dims[2], np.int32)
        self.R.ref_distmap_d2(self.h, _ip(out))
        return out.reshape(self.dims)

    def distance(self, x, y, z):
        return self.R.ref_distmap_distance(self.h, float(x), float(y), float(z))

    def world_to_grid(self, p):
        g = np.zeros(3, np.int32)
        self.R.ref_distmap_world_to_grid(self.h, float(p[0]), float(p[1]), float(p[2]), _ip(g))
        return g


_REF_ARA = None


def ref_arastar_lib():
    """The reference's own ARA* (smpl/src/search/arastar.cpp compiled from /root/reference), or None."""
    global _REF_ARA
    if _REF_ARA is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_arastar.so")
        if not os.path.exists(path):
            return None
        _REF_ARA = C.CDLL(path)
    return _REF_ARA


def arastar_search(which, off, dst, cost, h, start, goal, eps, max_expansions, max_path=4096):
    """ARA* (first solution at `eps`, at most `max_expansions` expansions) on an explicit CSR graph.
    which = "oracle" (oracle/arastar.h) or "reference" (the reference's arastar.cpp).
    Returns dict(found, cost, expansions, path)."""
    off = np.ascontiguousarray(off, dtype=np.int32)
    dst = np.ascontiguousarray(dst, dtype=np.int32)
    cost = np.ascontiguousarray(cost, dtype=np.int32)
    h = np.ascontiguousarray(h, dtype=np.int32)
    path = np.zeros(max_path, np.int32)
    out = np.zeros(4, np.int32)
    if which == "reference":
        fn = ref_arastar_lib().ref_arastar_search
    else:
        fn = lib().oracle_arastar_search
    fn(len(off) - 1, _ip(off), _ip(dst), _ip(cost), _ip(h), int(start), int(goal), C.c_double(eps), int(max_expansions),
       _ip(path), max_path, _ip(out))
    n = int(out[3])
    return dict(found=bool(out[0]), cost=int(out[1]), expansions=int(out[2]), path=path[:min(n, max_path)].copy())


_REF_SC = None


def ref_shortcut_lib():
    """The reference's own shortcut templates (smpl/geometry/shortcut.h instantiated by oracle/ref_shortcut_shim.cpp,
    compiled from /root/reference), or None."""
    global _REF_SC
    if _REF_SC is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_shortcut.so")
        if not os.path.exists(path):
            return None
        _REF_SC = C.CDLL(path)
    return _REF_SC


def shortcut_table(which, costs, valid, pair_cost, algo, granularity=1):
    """shortcut::ShortcutPath (algo 0) / DivideAndConquerShortcutPath (algo 1) on an index path of n points:
    costs[n-1] of the original segments, valid[n][n] and pair_cost[n][n] = the generator's answer for (i, j).
    which = "oracle" (oracle/shortcut.h) or "reference".  Returns the output indices, or None on failure."""
    costs = np.ascontiguousarray(costs, dtype=np.float64)
    n = len(costs) + 1 if len(valid) else 0
    valid = np.ascontiguousarray(valid, dtype=np.uint8).reshape(n, n)
    pair_cost = np.ascontiguousarray(pair_cost, dtype=np.float64).reshape(n, n)
    out = np.zeros(max(2 * n, 4), np.int32)
    if which == "reference":
        VF = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_void_p)
        CF = C.CFUNCTYPE(C.c_double, C.c_int, C.c_int, C.c_void_p)
        vf = VF(lambda a, b, u: int(valid[a, b]))
        cf = CF(lambda a, b, u: float(pair_cost[a, b]))
        r = ref_shortcut_lib().ref_shortcut_path(n, _dp(costs), vf, cf, None, int(algo), int(granularity), _ip(out), len(out))
    else:
        r = lib().oracle_shortcut_table(n, _dp(costs), _bp(valid), _dp(pair_cost), int(algo), int(granularity),
                                        _ip(out), len(out))
    return None if r < 0 else out[:r].copy()


_REF_VOX = None


def ref_voxelize_lib():
    """The reference's own voxeliser (smpl/src/geometry/voxelize.cpp + mesh_utils.cpp compiled from /root/reference
    against the arithmetic Eigen stand-in), or None."""
    global _REF_VOX
    if _REF_VOX is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_voxelize.so")
        if not os.path.exists(path):
            return None
        _REF_VOX = C.CDLL(path)
    return _REF_VOX


def _vox_lib(which):
    return (ref_voxelize_lib(), "ref_") if which == "reference" else (lib(), "oracle_")


def voxelize_mesh(which, vertices, triangles, res, voxel_origin=None, fill=False):
    """geometry::VoxelizeMesh: voxel centres [n][3] in ExtractVoxels order.  voxel_origin None = half-res grid."""
    L, pre = _vox_lib(which)
    v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
    o = None if voxel_origin is None else np.ascontiguousarray(voxel_origin, dtype=np.float64)
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 3), np.float64)
        n = getattr(L, pre + "voxelize_mesh")(_dp(v), len(v), _ip(t), len(t), C.c_double(res),
                                              None if o is None else _dp(o), int(fill), _dp(out), cap)
        if n >= 0:
            return out[:n].copy()
        cap = -n


def voxelize_box(which, size, pose3x4, res, voxel_origin=None, fill=False):
    """geometry::VoxelizeBox(length, width, height, pose, res[, voxel_origin], voxels, fill)."""
    L, pre = _vox_lib(which)
    p = np.ascontiguousarray(pose3x4, dtype=np.float64).reshape(3, 4)
    o = None if voxel_origin is None else np.ascontiguousarray(voxel_origin, dtype=np.float64)
    cap = 1 << 16
    while True:
        out = np.zeros((cap, 3), np.float64)
        n = getattr(L, pre + "voxelize_box")(C.c_double(size[0]), C.c_double(size[1]), C.c_double(size[2]), _dp(p),
                                             C.c_double(res), None if o is None else _dp(o), int(fill), _dp(out), cap)
        if n >= 0:
            return out[:n].copy()
        cap = -n


def shape_mesh(which, kind, dims):
    """geometry::CreateIndexed{Box,Sphere,Cylinder,Cone}Mesh (kind 0..3) at the origin: (vertices, triangles)."""
    L, pre = _vox_lib(which)
    d = np.zeros(3, np.float64)
    d[:len(dims)] = dims
    v = np.zeros((64, 3), np.float64)
    t = np.zeros(3 * 128, np.int32)
    ni = C.c_int(0)
    nv = getattr(L, pre + "shape_mesh")(int(kind), _dp(d), _dp(v), _ip(t), C.byref(ni))
    return v[:nv].copy(), t[:ni.value].reshape(-1, 3).copy()


def box_mesh(which, size):
    """geometry::CreateIndexedBoxMesh: (vertices[8][3], triangles[12][3])."""
    L, pre = _vox_lib(which)
    v = np.zeros((8, 3), np.float64)
    t = np.zeros(36, np.int32)
    getattr(L, pre + "box_mesh")(C.c_double(size[0]), C.c_double(size[1]), C.c_double(size[2]), _dp(v), _ip(t))
    return v, t.reshape(12, 3)


class OracleScene:
    def __init__(self, robot_path, group, planning_joints, origin, size, res, max_dist):
        L = lib()
        self.L = L
        o = np.asarray(origin, dtype=np.float64)
        s = np.asarray(size, dtype=np.float64)
        self.h = L.oracle_scene_create(robot_path.encode(), group.encode(), ",".join(planning_joints).encode(),
                                       _dp(o), _dp(s), float(res), float(max_dist))
        if not self.h:
            raise RuntimeError("oracle_scene_create: " + L.oracle_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.dof = len(planning_joints)

    def close(self):
        if self.h:
            self.L.oracle_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene ----
    def set_joint(self, name, value):
        return self.L.oracle_scene_set_joint(self.h, name.encode(), C.c_double(value))

    def use_desc_acm(self):
        self.L.oracle_scene_use_desc_acm(self.h)

    def acm_set(self, a, b, allowed):
        self.L.oracle_scene_acm_set(self.h, a.encode(), b.encode(), int(allowed))

    def set_padding(self, p):
        self.L.oracle_scene_set_padding(self.h, C.c_double(p))

    def add_cells(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 3)
        self.L.oracle_scene_add_cells(self.h, _ip(cells), len(cells))

    def add_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self.L.oracle_scene_add_points(self.h, _dp(pts), len(pts))

    def remove_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self.L.oracle_scene_remove_points(self.h, _dp(pts), len(pts))

    def attach_spheres(self, body_id, link, centers, radius):
        c = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
        r = self.L.oracle_scene_attach_spheres(self.h, body_id.encode(), link.encode(), _dp(c), len(c),
                                               C.c_double(radius))
        if r != 0:
            raise RuntimeError("attach failed")

    def prime(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        self.L.oracle_scene_prime(self.h, _dp(q))

    def grid_info(self):
        dims = np.zeros(3, np.int32)
        origin = np.zeros(3, np.float64)
        res = C.c_double()
        dmax = C.c_int32()
        self.L.oracle_scene_grid_info(self.h, _ip(dims), _dp(origin), C.byref(res), C.byref(dmax))
        return dims, origin, res.value, dmax.value

    def df_d2(self):
        dims, _, _, _ = self.grid_info()
        out = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.int32)
        self.L.oracle_scene_df_d2(self.h, _ip(out))
        return out.reshape(tuple(int(d) for d in dims))

    def world_to_grid(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((len(xyz), 3), np.int32)
        self.L.oracle_world_to_grid(self.h, _dp(xyz), len(xyz), _ip(out))
        return out

    # ---- validity ----
    def _q(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.dof)
        return q

    def is_states_valid(self, q):
        q = self._q(q)
        v = np.zeros(len(q), np.uint8)
        self.L.oracle_is_states_valid(self.h, _dp(q), len(q), _bp(v))
        return v

    def time_states_valid(self, q):
        q = self._q(q)
        v = np.zeros(len(q), np.uint8)
        t = self.L.oracle_time_states_valid(self.h, _dp(q), len(q), _bp(v))
        return t, v

    def report_states(self, q):
        q = self._q(q)
        n = len(q)
        v = np.zeros(n, np.uint8)
        L = np.zeros(n, np.int32)
        cm = np.zeros(n, np.float64)
        pm = np.zeros(n, np.float64)
        self.L.oracle_report_states(self.h, _dp(q), n, _bp(v), _ip(L), _dp(cm), _dp(pm))
        return v, L, cm, pm

    def is_edges_valid(self, q0, q1):
        q0 = self._q(q0)
        q1 = self._q(q1)
        v = np.zeros(len(q0), np.uint8)
        c = np.zeros(len(q0), np.int32)
        self.L.oracle_is_edges_valid(self.h, _dp(q0), _dp(q1), len(q0), _bp(v), _ip(c))
        return v, c

    def time_edges_valid(self, q0, q1):
        q0 = self._q(q0)
        q1 = self._q(q1)
        v = np.zeros(len(q0), np.uint8)
        c = np.zeros(len(q0), np.int32)
        t = self.L.oracle_time_edges_valid(self.h, _dp(q0), _dp(q1), len(q0), _bp(v), _ip(c))
        return t, v, c

    def report_edges(self, q0, q1):
        q0 = self._q(q0)
        q1 = self._q(q1)
        v = np.zeros(len(q0), np.uint8)
        c = np.zeros(len(q0), np.int32)
        L = np.zeros(len(q0), np.int32)
        self.L.oracle_report_edges(self.h, _dp(q0), _dp(q1), len(q0), _bp(v), _ip(c), _ip(L))
        return v, c, L

    def edge_waypoints(self, q0, q1, max_wp=256):
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        q1 = np.ascontiguousarray(q1, dtype=np.float64)
        out = np.zeros((max_wp, self.dof), np.float64)
        n = self.L.oracle_edge_waypoints(self.h, _dp(q0), _dp(q1), _dp(out), max_wp)
        return out[:n].copy()

    def num_nodes(self):
        return self.L.oracle_num_nodes(self.h)

    def sphere_centers(self, q):
        q = self._q(q)
        nn = self.num_nodes()
        out = np.zeros((len(q), nn, 3), np.float64)
        self.L.oracle_sphere_centers(self.h, _dp(q), len(q), _dp(out))
        return out

    def collision_distance(self, q):
        """CollisionSpace::collisionDistance per state."""
        q = self._q(q)
        out = np.zeros(len(q), np.float64)
        self.L.oracle_collision_distance(self.h, _dp(q), len(q), _dp(out))
        return out

    def node_table(self):
        nn = self.num_nodes()
        out = np.zeros((nn, 8), np.float64)
        n = self.L.oracle_node_table(self.h, _dp(out))
        assert n == nn
        return out

    def motion_weights(self):
        w = np.zeros(self.dof, np.float64)
        t = np.zeros(self.dof, np.int32)
        self.L.oracle_motion_weights(self.h, _dp(w), _ip(t))
        return w, t

    def checked_pairs(self):
        out = np.zeros((4096, 2), np.int32)
        n = self.L.oracle_checked_pairs(self.h, _ip(out), 4096)
        return out[:n].copy()

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.L.oracle_scene_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(df_lookups=a.value, pair_tests=b.value, link_updates=c.value)

    # ---- planning model / heuristic ----
    def init_kdl(self, chain_root, chain_tip, planning_link, T_kin_to_planning=None, xyz_offset=None):
        T = np.eye(4)[:3] if T_kin_to_planning is None else np.asarray(T_kin_to_planning, np.float64).reshape(3, 4)
        T = np.ascontiguousarray(T, dtype=np.float64)
        off = np.zeros(3) if xyz_offset is None else np.asarray(xyz_offset, np.float64)
        off = np.ascontiguousarray(off, dtype=np.float64)
        r = self.L.oracle_scene_init_kdl(self.h, chain_root.encode(), chain_tip.encode(), planning_link.encode(),
                                         _dp(T), _dp(off))
        if r != 0:
            raise RuntimeError("init_kdl: " + self.L.oracle_last_error().decode())

    def check_joint_limits(self, q):
        q = self._q(q)
        out = np.zeros(len(q), np.uint8)
        self.L.oracle_check_joint_limits(self.h, _dp(q), len(q), _bp(out))
        return out

    def joint_limits(self):
        lo = np.zeros(self.dof)
        hi = np.zeros(self.dof)
        c = np.zeros(self.dof, np.uint8)
        self.L.oracle_joint_limits(self.h, _dp(lo), _dp(hi), _bp(c))
        return lo, hi, c

    def planning_frame_fk(self, q):
        q = self._q(q)
        out = np.zeros((len(q), 6), np.float64)
        self.L.oracle_planning_frame_fk(self.h, _dp(q), len(q), _dp(out))
        return out

    def heur_init(self, inflation_radius, cost_per_cell):
        return self.L.oracle_heur_init(self.h, C.c_double(inflation_radius), int(cost_per_cell))

    def heur_set_goal(self, x, y, z):
        return self.L.oracle_heur_set_goal(self.h, C.c_double(x), C.c_double(y), C.c_double(z))

    def heur_grid(self):
        dims, _, _, _ = self.grid_info()
        n = int(dims[0] + 2) * int(dims[1] + 2) * int(dims[2] + 2)
        out = np.zeros(n, np.int32)
        self.L.oracle_heur_grid(self.h, _ip(out))
        return out.reshape(int(dims[2] + 2), int(dims[1] + 2), int(dims[0] + 2))

    def goal_heuristics(self, q):
        q = self._q(q)
        h = np.zeros(len(q), np.int32)
        self.L.oracle_goal_heuristics(self.h, _dp(q), len(q), _ip(h))
        return h


    def attach_box(self, body_id, link, size, pose3x4):
        """attachBody for a box shape: spheres generated from the shape's surface voxels."""
        sz = np.ascontiguousarray(size, dtype=np.float64)
        p = np.ascontiguousarray(pose3x4, dtype=np.float64).reshape(3, 4)
        return self.L.oracle_scene_attach_box(self.h, body_id.encode(), link.encode(), _dp(sz), _dp(p))

    def detach(self, body_id):
        return self.L.oracle_scene_detach(self.h, body_id.encode())

    def insert_boxes(self, boxes):
        """WorldCollisionModel::insertObject for box primitives: boxes[n][15] = size(3), pose 3x4 row-major."""
        b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 15)
        return self.L.oracle_scene_insert_boxes(self.h, _dp(b), len(b))

    def shortcut_path(self, path, continuous, kind=0):
        """ShortcutPath(rm, cc, pin, pout, type) (post_processing.cpp:284-365): indices of the shortcut path's
        points and the number of isStateToStateValid calls made.  kind 0 = JOINT_SPACE, 1 = JOINT_POSITION_VELOCITY_SPACE."""
        path = self._q(path)
        cont = np.ascontiguousarray(continuous, dtype=np.uint8)
        out = np.zeros(max(len(path), 1), np.int32)
        checks = C.c_int64(0)
        n = self.L.oracle_shortcut_path(self.h, _dp(path), len(path), _bp(cont), int(kind), _ip(out), C.byref(checks))
        return out[:n].copy(), int(checks.value)

    def interpolate_path(self, path, max_points=1 << 16):
        """InterpolatePath (post_processing.cpp:476-540)."""
        path = self._q(path)
        out = np.zeros((max_points, path.shape[1]), np.float64)
        n = self.L.oracle_interpolate_path(self.h, _dp(path), len(path), _dp(out), max_points)
        if n < 0:
            raise RuntimeError("interpolate_path: more than %d points" % max_points)
        return out[:n].copy()

    def plan(self, start, goal_xyz, params, max_path=4096, lazy=False):
        """params: smpl_b200.scenes.PlanParams.  Returns dict(success, expansions, cost, path_ids, num_states, seconds);
        lazy: the lazy successors under oracle/lazy_arastar.h (+ `evaluations`)."""
        start = np.ascontiguousarray(start, dtype=np.float64)
        goal = np.ascontiguousarray(goal_xyz, dtype=np.float64)
        res = np.ascontiguousarray(params.resolutions, dtype=np.float64)
        prims = np.ascontiguousarray(params.mprims, dtype=np.float64)
        flags = np.ascontiguousarray(params.short_flags, dtype=np.uint8)
        tol = np.ascontiguousarray(params.xyz_tolerance, dtype=np.float64)
        summary = np.zeros(8, np.int32)
        path = np.zeros(max_path, np.int32)
        pstates = np.zeros((max_path, len(start)), np.float64)
        w = getattr(params, "weights", None)
        w = np.zeros(0) if w is None else np.ascontiguousarray(w, dtype=np.float64)
        self.L.oracle_set_prim_weights(_dp(w), len(w))
        fn = self.L.oracle_plan_lazy if lazy else self.L.oracle_plan
        secs = fn(self.h, _dp(start), _dp(goal), _dp(res), _dp(prims), _bp(flags), len(prims),
                                  int(params.use_short_dist), C.c_double(params.short_dist_thresh),
                                  C.c_double(params.epsilon), int(params.max_expansions), _dp(tol),
                                  _ip(summary), _ip(path), max_path, _dp(pstates))
        n = int(summary[3])
        assert int(summary[5]) == n, "extractPath failed"
        out = dict(path_states=pstates[:min(n, max_path)].copy(), success=bool(summary[0]), expansions=int(summary[1]), cost=int(summary[2]),
                   path_ids=path[:min(n, max_path)].copy(), num_states=int(summary[4]), seconds=float(secs))
        if lazy:
            out["evaluations"] = int(self.L.oracle_last_lazy_evaluations())
        return out


_REF_CC = None


def ref_collision_lib():
    """The reference's own collision checker (oracle/_ref/libref_collision.so, see oracle/ref_collision_shim.cpp), or None."""
    global _REF_CC
    if _REF_CC is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_collision.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.refcc_last_error.restype = C.c_char_p
        R.refcc_create.restype = C.c_void_p
        R.refcc_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, c_double_p, c_double_p, C.c_double, C.c_double]
        R.refcc_time_states_valid.restype = C.c_double
        R.refcc_time_edges_valid.restype = C.c_double
        _REF_CC = R
    return _REF_CC


class RefCollisionScene:
    """Same driving surface as OracleScene, served by the reference's CollisionSpace."""

    def __init__(self, robot_path, group, planning_joints, origin, size, res, max_dist):
        R = ref_collision_lib()
        self.R = R
        o = np.asarray(origin, dtype=np.float64)
        s = np.asarray(size, dtype=np.float64)
        self.h = R.refcc_create(robot_path.encode(), group.encode(), ",".join(planning_joints).encode(),
                                _dp(o), _dp(s), float(res), float(max_dist))
        if not self.h:
            raise RuntimeError("refcc_create: " + R.refcc_last_error().decode())
        self.h = C.c_void_p(self.h)
        self.dof = len(planning_joints)

    def close(self):
        if self.h:
            self.R.refcc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_joint(self, name, value):
        return self.R.refcc_set_joint(self.h, name.encode(), C.c_double(value))

    def use_desc_acm(self):
        self.R.refcc_use_desc_acm(self.h)

    def acm_set(self, a, b, allowed):
        self.R.refcc_acm_set(self.h, a.encode(), b.encode(), int(allowed))

    def set_padding(self, p):
        self.R.refcc_set_padding(self.h, C.c_double(p))

    def add_cells(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1, 3)
        self.R.refcc_add_cells(self.h, _ip(cells), len(cells))

    def add_points(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        self.R.refcc_add_points(self.h, _dp(pts), len(pts))

    def insert_boxes(self, boxes):
        b = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 15)
        if self.R.refcc_insert_boxes(self.h, _dp(b), len(b)) != 0:
            raise RuntimeError("insertObject failed")

    def insert_shapes(self, rows):
        """rows[n][16] = kind (0 box, 1 sphere, 2 cylinder, 3 cone), 3 dimensions, pose 3x4: insertObject per shape"""
        r = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 16)
        if self.R.refcc_insert_shapes(self.h, _dp(r), len(r)) != 0:
            raise RuntimeError("insertObject failed")

    def attach_box(self, body_id, link, size, pose3x4):
        sz = np.ascontiguousarray(size, dtype=np.float64)
        p = np.ascontiguousarray(pose3x4, dtype=np.float64).reshape(3, 4)
        n = self.R.refcc_attach_box(self.h, body_id.encode(), link.encode(), _dp(sz), _dp(p))
        if n < 0:
            raise RuntimeError("attachObject failed")
        return n

    def detach(self, body_id):
        return self.R.refcc_detach(self.h, body_id.encode())

    def prime(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        self.R.refcc_prime(self.h, _dp(q))

    def df_d2(self):
        dims = np.zeros(3, np.int32)
        self.R.refcc_grid_dims(self.h, _ip(dims))
        out = np.zeros(int(dims[0]) * int(dims[1]) * int(dims[2]), np.int32)
        self.R.refcc_df_d2(self.h, _ip(out))
        return out.reshape(tuple(int(d) for d in dims))

    def _q(self, q):
        return np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.dof)

    def is_states_valid(self, q):
        q = self._q(q)
        v = np.zeros(len(q), np.uint8)
        self.R.refcc_is_states_valid(self.h, _dp(q), len(q), _bp(v))
        return v

    def is_edges_valid(self, q0, q1):
        q0 = self._q(q0)
        q1 = self._q(q1)
        v = np.zeros(len(q0), np.uint8)
        c = np.zeros(len(q0), np.int32)
        self.R.refcc_is_edges_valid(self.h, _dp(q0), _dp(q1), len(q0), _bp(v), _ip(c))
        return v, c

    def time_states_valid(self, q):
        q = self._q(q)
        v = np.zeros(len(q), np.uint8)
        t = self.R.refcc_time_states_valid(self.h, _dp(q), len(q), _bp(v))
        return t, v

    def time_edges_valid(self, q0, q1):
        q0 = self._q(q0)
        q1 = self._q(q1)
        v = np.zeros(len(q0), np.uint8)
        c = np.zeros(len(q0), np.int32)
        t = self.R.refcc_time_edges_valid(self.h, _dp(q0), _dp(q1), len(q0), _bp(v), _ip(c))
        return t, v, c

    def edge_waypoints(self, q0, q1, max_wp=256):
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        q1 = np.ascontiguousarray(q1, dtype=np.float64)
        out = np.zeros((max_wp, self.dof), np.float64)
        n = self.R.refcc_edge_waypoints(self.h, _dp(q0), _dp(q1), _dp(out), max_wp)
        return out[:n].copy()

    def collision_distance(self, q):
        """the reference's own CollisionSpace::collisionDistance per state."""
        q = self._q(q)
        out = np.zeros(len(q), np.float64)
        self.R.refcc_collision_distance(self.h, _dp(q), len(q), _dp(out))
        return out

    def node_table(self, max_nodes=4096):
        out = np.zeros((max_nodes, 8), np.float64)
        n = self.R.refcc_node_table(self.h, _dp(out))
        return out[:n].copy()

    def sphere_centers(self, q, n_nodes):
        q = self._q(q)
        out = np.zeros((len(q), n_nodes, 3), np.float64)
        self.R.refcc_sphere_centers(self.h, _dp(q), len(q), _dp(out))
        return out

    def motion_weights(self):
        w = np.zeros(self.dof, np.float64)
        self.R.refcc_motion_weights(self.h, _dp(w))
        return w

    def limits(self):
        lo = np.zeros(self.dof)
        hi = np.zeros(self.dof)
        c = np.zeros(self.dof, np.uint8)
        self.R.refcc_limits(self.h, _dp(lo), _dp(hi), _bp(c))
        return lo, hi, c

    def kdl_fk_and_limits(self, scene, q):
        """KDLRobotModel::computePlanningLinkFK + checkJointLimits of the reference's own kdl_robot_model.cpp (over the
        KDL stand-in): (pose6[n][6], within[n])"""
        q = self._q(q)
        T = np.ascontiguousarray(np.asarray(scene.T_kin_to_planning, np.float64).reshape(3, 4))
        pose = np.zeros((len(q), 6), np.float64)
        ok = np.zeros(len(q), np.uint8)
        rc = self.R.refcc_kdl_fk_and_limits(self.h, scene.chain_root.encode(), scene.chain_tip.encode(),
                                            scene.planning_link.encode(), _dp(T), _dp(q), len(q), _dp(pose), _bp(ok))
        if rc != 0:
            raise RuntimeError("refcc_kdl_fk_and_limits: %d" % rc)
        return pose, ok

    def goal_heuristics(self, scene, goal_xyz, resolutions, q):
        """BfsHeuristic::GetGoalHeuristic of the reference for joint states q (by lattice state id) -> (h[n], h of the
        goal state, getMetricGoalDistance[n])"""
        q = self._q(q)
        T = np.ascontiguousarray(np.asarray(scene.T_kin_to_planning, np.float64).reshape(3, 4))
        off = np.ascontiguousarray(scene.xyz_offset, dtype=np.float64)
        goal = np.ascontiguousarray(goal_xyz, dtype=np.float64)
        res = np.ascontiguousarray(resolutions, dtype=np.float64)
        h = np.zeros(len(q) + 1, np.int32)
        metric = np.zeros(len(q), np.float64)
        rc = self.R.refcc_goal_heuristics(self.h, scene.chain_root.encode(), scene.chain_tip.encode(),
                                          scene.planning_link.encode(), _dp(T), _dp(off),
                                          C.c_double(scene.inflation_radius), int(scene.cost_per_cell), _dp(goal), _dp(res),
                                          _dp(q), len(q), _ip(h), _dp(metric))
        if rc != 0:
            raise RuntimeError("refcc_goal_heuristics: %d" % rc)
        return h[:-1].copy(), int(h[-1]), metric

    def kdl_limits(self, scene):
        lo = np.zeros(self.dof)
        hi = np.zeros(self.dof)
        c = np.zeros(self.dof, np.uint8)
        if self.R.refcc_kdl_limits(self.h, scene.chain_root.encode(), scene.chain_tip.encode(), _dp(lo), _dp(hi), _bp(c)) != 0:
            raise RuntimeError("refcc_kdl_limits")
        return lo, hi, c

    def post_process(self, scene, path, kind, max_points=1 << 16):
        """ShortcutPath (kind 0 JOINT_SPACE, 1 JOINT_POSITION_VELOCITY_SPACE) / InterpolatePath (kind 2) of the
        reference's post_processing.cpp; returns the points, or None when the reference reports failure."""
        path = self._q(path)
        out = np.zeros((max_points, self.dof), np.float64)
        n = self.R.refcc_post_process(self.h, scene.chain_root.encode(), scene.chain_tip.encode(),
                                      scene.planning_link.encode(), _dp(path), len(path), int(kind), _dp(out), max_points)
        if n == -3:
            return None
        if n < 0:
            raise RuntimeError("refcc_post_process: %d" % n)
        return out[:n].copy()

    def plan(self, scene, start, goal_xyz, params, max_path=4096, lazy=False):
        """One query through the reference's ManipLattice + BfsHeuristic + ARAStar (oracle/ref_planner_shim.cpp);
        same result dict as OracleScene.plan.  lazy: GetLazySuccs / GetTrueCost under the reference's LazyARAStar
        instead (the dict then also holds `evaluations`, the GetTrueCost calls)."""
        start = np.ascontiguousarray(start, dtype=np.float64)
        goal = np.ascontiguousarray(goal_xyz, dtype=np.float64)
        res = np.ascontiguousarray(params.resolutions, dtype=np.float64)
        prims = np.ascontiguousarray(params.mprims, dtype=np.float64)
        flags = np.ascontiguousarray(params.short_flags, dtype=np.uint8)
        tol = np.ascontiguousarray(params.xyz_tolerance, dtype=np.float64)
        T = np.ascontiguousarray(np.asarray(scene.T_kin_to_planning, np.float64).reshape(3, 4))
        off = np.ascontiguousarray(scene.xyz_offset, dtype=np.float64)
        summary = np.zeros(8, np.int32)
        path = np.zeros(max_path, np.int32)
        pstates = np.zeros((max_path, len(start)), np.float64)
        w = getattr(params, "weights", None)
        w = np.zeros(0) if w is None else np.ascontiguousarray(w, dtype=np.float64)
        self.R.refcc_set_prim_weights(_dp(w), len(w))
        fn = self.R.refcc_plan_lazy if lazy else self.R.refcc_plan
        rc = fn(self.h, scene.chain_root.encode(), scene.chain_tip.encode(), scene.planning_link.encode(),
                               _dp(T), _dp(off), C.c_double(scene.inflation_radius), int(scene.cost_per_cell),
                               _dp(start), _dp(goal), _dp(res), _dp(prims), _bp(flags), len(prims),
                               int(params.use_short_dist), C.c_double(params.short_dist_thresh),
                               C.c_double(params.epsilon), int(params.max_expansions), _dp(tol),
                               _ip(summary), _ip(path), max_path, _dp(pstates))
        if rc != 0:
            raise RuntimeError("refcc_plan: the reference refused step %d" % -rc)
        n = int(summary[3])
        out = dict(path_states=pstates[:min(int(summary[5]), max_path)].copy(), success=bool(summary[0]),
                   expansions=int(summary[1]), cost=int(summary[2]), path_ids=path[:min(n, max_path)].copy(),
                   num_states=int(summary[4]))
        if lazy:
            out["evaluations"] = int(self.R.refcc_last_lazy_evaluations())
        return out


class _BfsBase:
    prefix = None

    def __init__(self, L, nx, ny, nz):
        self.L = L
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.h = C.c_void_p(getattr(L, self.prefix + "_create")(self.nx, self.ny, self.nz))

    def close(self):
        if self.h:
            getattr(self.L, self.prefix + "_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_walls(self, walls_zyx):
        """walls_zyx: uint8 array [nz, ny, nx] (x fastest), nonzero = wall."""
        w = np.ascontiguousarray(walls_zyx, dtype=np.uint8)
        assert w.shape == (self.nz, self.ny, self.nx)
        getattr(self.L, self.prefix + "_set_walls")(self.h, _bp(w))

    def run(self, x, y, z):
        return getattr(self.L, self.prefix + "_run")(self.h, int(x), int(y), int(z))

    def run_multi(self, seeds):
        s = np.ascontiguousarray(seeds, dtype=np.int32).reshape(-1, 3)
        return getattr(self.L, self.prefix + "_run_multi")(self.h, _ip(s), len(s))

    def grid(self):
        out = np.zeros((self.nz + 2) * (self.ny + 2) * (self.nx + 2), np.int32)
        getattr(self.L, self.prefix + "_grid")(self.h, _ip(out))
        return out.reshape(self.nz + 2, self.ny + 2, self.nx + 2)


class OracleBfs(_BfsBase):
    prefix = "oracle_bfs"

    def __init__(self, nx, ny, nz):
        super().__init__(lib(), nx, ny, nz)

    def time_run(self, x, y, z):
        return self.L.oracle_time_bfs_run(self.h, int(x), int(y), int(z))


class RefBfs(_BfsBase):
    """The reference's own BFS_3D class (oracle/_ref)."""
    prefix = "ref_bfs"

    def __init__(self, nx, ny, nz):
        R = ref_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libref_bfs3d.so not built")
        super().__init__(R, nx, ny, nz)

    def time_run(self, x, y, z):
        return self.L.ref_bfs_time_run(self.h, int(x), int(y), int(z))
