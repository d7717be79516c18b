// ORACLE (test infrastructure only).  Stand-in for orocos_kdl (absent, un-vendored third-party dependency of the
// reference's sbpl_kdl_robot_model; version unpinned, ROS-Indigo era = KDL 1.3.x): the classes and functions
// kdl_robot_model.cpp names, with the arithmetic restated from KDL's published sources (frames.inl, frames.cpp,
// joint.cpp, segment.cpp, chainfksolverpos_recursive.cpp, tree.cpp) in their operation order -- the same restatement
// as oracle/kdl_model.cpp.  What compiling the reference's own kdl_robot_model.cpp against it pins is that file's
// control flow (which segment number it asks for, angle normalisation, limit selection and tests); KDL's own
// arithmetic stays restated.  The inverse-kinematics solvers are inert (never on the hot path).
#pragma once

#include <cmath>
#include <map>
#include <string>
#include <vector>

namespace KDL {

inline double sqr(double x) { return x * x; }

class Vector
{
public:
    double data[3];
    Vector() { data[0] = data[1] = data[2] = 0.0; }
    Vector(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int i) const { return data[i]; }
    double& operator[](int i) { return data[i]; }
    double operator()(int i) const { return data[i]; }
    double& operator()(int i) { return data[i]; }
    double x() const { return data[0]; }
    double y() const { return data[1]; }
    double z() const { return data[2]; }
    void x(double v) { data[0] = v; }
    void y(double v) { data[1] = v; }
    void z(double v) { data[2] = v; }
    static Vector Zero() { return Vector(); }
    double Norm() const   // frames.cpp: scaled to avoid overflow
    {
        double tmp1 = std::fabs(data[0]), tmp2 = std::fabs(data[1]);
        if (tmp1 >= tmp2) {
            tmp2 = std::fabs(data[2]);
            if (tmp1 >= tmp2) {
                if (tmp1 == 0) return 0;
                return tmp1 * std::sqrt(1 + sqr(data[1] / data[0]) + sqr(data[2] / data[0]));
            }
            return tmp2 * std::sqrt(1 + sqr(data[0] / data[2]) + sqr(data[1] / data[2]));
        }
        tmp1 = std::fabs(data[2]);
        if (tmp2 > tmp1) {
            return tmp2 * std::sqrt(1 + sqr(data[0] / data[1]) + sqr(data[2] / data[1]));
        }
        return tmp1 * std::sqrt(1 + sqr(data[0] / data[2]) + sqr(data[1] / data[2]));
    }
};
inline Vector operator+(const Vector& a, const Vector& b) { return Vector(a.data[0] + b.data[0], a.data[1] + b.data[1], a.data[2] + b.data[2]); }
inline Vector operator-(const Vector& a) { return Vector(-a.data[0], -a.data[1], -a.data[2]); }
inline Vector operator*(const Vector& a, double s) { return Vector(a.data[0] * s, a.data[1] * s, a.data[2] * s); }
inline Vector operator/(const Vector& a, double s) { return Vector(a.data[0] / s, a.data[1] / s, a.data[2] / s); }

class Rotation
{
public:
    double data[9];
    Rotation() { for (int i = 0; i < 9; ++i) data[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    Rotation(double Xx, double Yx, double Zx, double Xy, double Yy, double Zy, double Xz, double Yz, double Zz)
    {
        data[0] = Xx; data[1] = Yx; data[2] = Zx; data[3] = Xy; data[4] = Yy; data[5] = Zy; data[6] = Xz; data[7] = Yz; data[8] = Zz;
    }
    static Rotation Identity() { return Rotation(); }
    double operator()(int i, int j) const { return data[3 * i + j]; }
    Vector operator*(const Vector& v) const
    {
        return Vector(data[0] * v.data[0] + data[1] * v.data[1] + data[2] * v.data[2],
                      data[3] * v.data[0] + data[4] * v.data[1] + data[5] * v.data[2],
                      data[6] * v.data[0] + data[7] * v.data[1] + data[8] * v.data[2]);
    }
    Rotation Inverse() const { return Rotation(data[0], data[3], data[6], data[1], data[4], data[7], data[2], data[5], data[8]); }
    Vector Inverse(const Vector& v) const
    {
        return Vector(data[0] * v.data[0] + data[3] * v.data[1] + data[6] * v.data[2],
                      data[1] * v.data[0] + data[4] * v.data[1] + data[7] * v.data[2],
                      data[2] * v.data[0] + data[5] * v.data[1] + data[8] * v.data[2]);
    }
    static Rotation Rot2(const Vector& v, double angle)   // frames.cpp
    {
        const double ct = std::cos(angle), st = std::sin(angle), vt = 1 - ct;
        const double m_vt_0 = vt * v(0), m_vt_1 = vt * v(1), m_vt_2 = vt * v(2);
        const double m_st_0 = v(0) * st, m_st_1 = v(1) * st, m_st_2 = v(2) * st;
        const double m_vt_0_1 = m_vt_0 * v(1), m_vt_0_2 = m_vt_0 * v(2), m_vt_1_2 = m_vt_1 * v(2);
        return Rotation(ct + m_vt_0 * v(0), -m_st_2 + m_vt_0_1, m_st_1 + m_vt_0_2,
                        m_st_2 + m_vt_0_1, ct + m_vt_1 * v(1), -m_st_0 + m_vt_1_2,
                        -m_st_1 + m_vt_0_2, m_st_0 + m_vt_1_2, ct + m_vt_2 * v(2));
    }
    static Rotation Quaternion(double x, double y, double z, double w)
    {
        const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
        return Rotation(w2 + x2 - y2 - z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
                        2 * x * y + 2 * w * z, w2 - x2 + y2 - z2, 2 * y * z - 2 * w * x,
                        2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, w2 - x2 - y2 + z2);
    }
    static Rotation RPY(double roll, double pitch, double yaw)
    {
        const double ca1 = std::cos(yaw), sa1 = std::sin(yaw), cb1 = std::cos(pitch), sb1 = std::sin(pitch);
        const double cc1 = std::cos(roll), sc1 = std::sin(roll);
        return Rotation(ca1 * cb1, ca1 * sb1 * sc1 - sa1 * cc1, ca1 * sb1 * cc1 + sa1 * sc1,
                        sa1 * cb1, sa1 * sb1 * sc1 + ca1 * cc1, sa1 * sb1 * cc1 - ca1 * sc1,
                        -sb1, cb1 * sc1, cb1 * cc1);
    }
    void GetRPY(double& roll, double& pitch, double& yaw) const
    {
        const double epsilon = 1E-12;
        pitch = std::atan2(-data[6], std::sqrt(sqr(data[0]) + sqr(data[3])));
        if (std::fabs(pitch) > (M_PI / 2.0 - epsilon)) {
            yaw = std::atan2(-data[1], data[4]);
            roll = 0.0;
        } else {
            roll = std::atan2(data[7], data[8]);
            yaw = std::atan2(data[3], data[0]);
        }
    }
};
inline Rotation operator*(const Rotation& lhs, const Rotation& rhs)
{
    return Rotation(
        lhs.data[0] * rhs.data[0] + lhs.data[1] * rhs.data[3] + lhs.data[2] * rhs.data[6],
        lhs.data[0] * rhs.data[1] + lhs.data[1] * rhs.data[4] + lhs.data[2] * rhs.data[7],
        lhs.data[0] * rhs.data[2] + lhs.data[1] * rhs.data[5] + lhs.data[2] * rhs.data[8],
        lhs.data[3] * rhs.data[0] + lhs.data[4] * rhs.data[3] + lhs.data[5] * rhs.data[6],
        lhs.data[3] * rhs.data[1] + lhs.data[4] * rhs.data[4] + lhs.data[5] * rhs.data[7],
        lhs.data[3] * rhs.data[2] + lhs.data[4] * rhs.data[5] + lhs.data[5] * rhs.data[8],
        lhs.data[6] * rhs.data[0] + lhs.data[7] * rhs.data[3] + lhs.data[8] * rhs.data[6],
        lhs.data[6] * rhs.data[1] + lhs.data[7] * rhs.data[4] + lhs.data[8] * rhs.data[7],
        lhs.data[6] * rhs.data[2] + lhs.data[7] * rhs.data[5] + lhs.data[8] * rhs.data[8]);
}

class Frame
{
public:
    Vector p;
    Rotation M;
    Frame() { }
    Frame(const Rotation& R, const Vector& V) : p(V), M(R) { }
    explicit Frame(const Vector& V) : p(V) { }
    explicit Frame(const Rotation& R) : M(R) { }
    static Frame Identity() { return Frame(); }
    Frame Inverse() const { return Frame(M.Inverse(), -M.Inverse(p)); }
};
inline Frame operator*(const Frame& lhs, const Frame& rhs) { return Frame(lhs.M * rhs.M, lhs.M * rhs.p + lhs.p); }

class JntArray
{
public:
    std::vector<double> data;
    JntArray() { }
    explicit JntArray(unsigned int n) : data(n, 0.0) { }
    void resize(unsigned int n) { data.assign(n, 0.0); }
    unsigned int rows() const { return (unsigned int)data.size(); }
    double operator()(unsigned int i) const { return data[i]; }
    double& operator()(unsigned int i) { return data[i]; }
};

class Joint
{
public:
    enum JointType { RotAxis, RotX, RotY, RotZ, TransAxis, TransX, TransY, TransZ, None };
    Joint() : type(None) { }
    Joint(const std::string& n, JointType t = None) : name(n), type(t) { }
    Joint(const std::string& n, const Vector& o, const Vector& a, JointType t) :   // joint.cpp: the axis is normalised
        name(n), type(t), axis(a / a.Norm()), origin(o) { }
    Frame pose(double q) const
    {
        switch (type) {
        case RotAxis: return Frame(Rotation::Rot2(axis, q), origin);
        case TransAxis: return Frame(origin + (axis * q));
        default: return Frame::Identity();
        }
    }
    const std::string& getName() const { return name; }
    JointType getType() const { return type; }
private:
    std::string name;
    JointType type;
    Vector axis, origin;
};

class Segment
{
public:
    Segment() { }
    Segment(const std::string& n, const Joint& j, const Frame& f) : name(n), joint(j), f_tip(j.pose(0).Inverse() * f) { }
    Frame pose(double q) const { return joint.pose(q) * f_tip; }
    const std::string& getName() const { return name; }
    const Joint& getJoint() const { return joint; }
private:
    std::string name;
    Joint joint;
    Frame f_tip;
};

class Chain
{
public:
    std::vector<Segment> segments;
    unsigned int nr_joints = 0;
    void addSegment(const Segment& s) { segments.push_back(s); if (s.getJoint().getType() != Joint::None) ++nr_joints; }
    unsigned int getNrOfSegments() const { return (unsigned int)segments.size(); }
    unsigned int getNrOfJoints() const { return nr_joints; }
    const Segment& getSegment(unsigned int i) const { return segments[i]; }
};

struct TreeElement { Segment segment; std::string parent; };
typedef std::map<std::string, TreeElement> SegmentMap;

class Tree
{
public:
    explicit Tree(const std::string& root = "root") : root_name(root) { }
    bool addSegment(const Segment& s, const std::string& hook_name)
    {
        if (hook_name != root_name && segments.find(hook_name) == segments.end()) return false;
        if (segments.find(s.getName()) != segments.end()) return false;
        TreeElement e;
        e.segment = s;
        e.parent = hook_name;
        segments[s.getName()] = e;
        return true;
    }
    const SegmentMap& getSegments() const { return segments; }
    /// tree.cpp getChain for chain_tip below chain_root (the only shape the reference's configurations use)
    bool getChain(const std::string& chain_root, const std::string& chain_tip, Chain& chain) const
    {
        chain = Chain();
        std::vector<const TreeElement*> up;
        std::string link = chain_tip;
        while (link != chain_root) {
            auto it = segments.find(link);
            if (it == segments.end()) return false;
            up.push_back(&it->second);
            link = it->second.parent;
        }
        for (auto it = up.rbegin(); it != up.rend(); ++it) chain.addSegment((*it)->segment);
        return true;
    }
private:
    std::string root_name;
    SegmentMap segments;
};

class ChainFkSolverPos_recursive
{
public:
    explicit ChainFkSolverPos_recursive(const Chain& c) : chain(c) { }
    int JntToCart(const JntArray& q_in, Frame& p_out, int segmentNr = -1)   // chainfksolverpos_recursive.cpp
    {
        if (segmentNr < 0) segmentNr = (int)chain.getNrOfSegments();
        p_out = Frame::Identity();
        if (q_in.rows() != chain.getNrOfJoints()) return -1;
        if (segmentNr > (int)chain.getNrOfSegments()) return -1;
        int j = 0;
        for (int i = 0; i < segmentNr; ++i) {
            if (chain.getSegment(i).getJoint().getType() != Joint::None) {
                p_out = p_out * chain.getSegment(i).pose(q_in(j));
                ++j;
            } else {
                p_out = p_out * chain.getSegment(i).pose(0.0);
            }
        }
        return 0;
    }
private:
    Chain chain;
};

class ChainIkSolverVel_pinv { public: explicit ChainIkSolverVel_pinv(const Chain&) { } };
class ChainIkSolverPos_NR_JL
{
public:
    ChainIkSolverPos_NR_JL(const Chain&, const JntArray&, const JntArray&, ChainFkSolverPos_recursive&, ChainIkSolverVel_pinv&, unsigned int, double) { }
    int CartToJnt(const JntArray&, const Frame&, JntArray&) { return -1; }   // inert: IK is not on the hot path
};

} // namespace KDL
