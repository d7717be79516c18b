// ORACLE (test infrastructure only).  Stand-in for an absent third-party header: just enough surface for the
// reference's sbpl_collision_checking sources to compile where they lie (see oracle/Makefile, target ref).
#pragma once
// urdfdom_headers data records (public fields only).  Rotation::setFromRPY / getQuaternion follow urdfdom's
// published pose.h; ModelInterface keeps links_/joints_ in name-ordered maps like urdfdom, and a link's
// child_joints / child_links in the order the model builder adds them (see ref_collision_shim.cpp).
#include <boost/shared_ptr.hpp>
#include <cmath>
#include <map>
#include <string>
#include <vector>
namespace urdf {
struct Vector3 { double x, y, z; Vector3(double x_ = 0.0, double y_ = 0.0, double z_ = 0.0) : x(x_), y(y_), z(z_) { } };
struct Rotation
{
    double x, y, z, w;
    Rotation(double x_ = 0.0, double y_ = 0.0, double z_ = 0.0, double w_ = 1.0) : x(x_), y(y_), z(z_), w(w_) { }
    void getQuaternion(double& qx, double& qy, double& qz, double& qw) const { qx = x; qy = y; qz = z; qw = w; }
    void setFromRPY(double roll, double pitch, double yaw)
    {
        const double phi = roll / 2.0, the = pitch / 2.0, psi = yaw / 2.0;
        x = sin(phi) * cos(the) * cos(psi) - cos(phi) * sin(the) * sin(psi);
        y = cos(phi) * sin(the) * cos(psi) + sin(phi) * cos(the) * sin(psi);
        z = cos(phi) * cos(the) * sin(psi) - sin(phi) * sin(the) * cos(psi);
        w = cos(phi) * cos(the) * cos(psi) + sin(phi) * sin(the) * sin(psi);
        normalize();
    }
    void normalize()
    {
        const double s = sqrt(x * x + y * y + z * z + w * w);
        if (s == 0.0) { x = 0.0; y = 0.0; z = 0.0; w = 1.0; }
        else { x /= s; y /= s; z /= s; w /= s; }
    }
};
struct Pose { Vector3 position; Rotation rotation; };
struct JointLimits { double lower = 0, upper = 0, effort = 0, velocity = 0; };
struct JointSafety { double soft_upper_limit = 0, soft_lower_limit = 0, k_position = 0, k_velocity = 0; };
struct Joint
{
    enum { UNKNOWN, REVOLUTE, CONTINUOUS, PRISMATIC, FLOATING, PLANAR, FIXED };
    std::string name;
    int type = UNKNOWN;
    Vector3 axis;
    std::string child_link_name, parent_link_name;
    Pose parent_to_joint_origin_transform;
    boost::shared_ptr<JointLimits> limits;
    boost::shared_ptr<JointSafety> safety;
};
struct Geometry { enum { SPHERE, BOX, CYLINDER, MESH } type; virtual ~Geometry() { } };
struct Sphere : Geometry { double radius = 0; Sphere() { type = SPHERE; } };
struct Box : Geometry { Vector3 dim; Box() { type = BOX; } };
struct Cylinder : Geometry { double length = 0, radius = 0; Cylinder() { type = CYLINDER; } };
struct Mesh : Geometry { std::string filename; Vector3 scale; Mesh() : scale(1, 1, 1) { type = MESH; } };
struct Collision { Pose origin; boost::shared_ptr<Geometry> geometry; std::string name; };
struct Link
{
    std::string name;
    boost::shared_ptr<Collision> collision;
    std::vector<boost::shared_ptr<Collision>> collision_array;
    boost::shared_ptr<Joint> parent_joint;
    std::vector<boost::shared_ptr<Joint>> child_joints;
    std::vector<boost::shared_ptr<Link>> child_links;
    boost::weak_ptr<Link> parent_link_;
    boost::shared_ptr<Link> getParent() const { return parent_link_.lock(); }
};
class ModelInterface
{
public:
    boost::shared_ptr<const Link> getRoot() const { return root_link_; }
    boost::shared_ptr<const Link> getLink(const std::string& name) const
    {
        auto it = links_.find(name);
        return it == links_.end() ? boost::shared_ptr<const Link>() : boost::shared_ptr<const Link>(it->second);
    }
    boost::shared_ptr<const Joint> getJoint(const std::string& name) const
    {
        auto it = joints_.find(name);
        return it == joints_.end() ? boost::shared_ptr<const Joint>() : boost::shared_ptr<const Joint>(it->second);
    }
    const std::string& getName() const { return name_; }
    std::map<std::string, boost::shared_ptr<Link>> links_;
    std::map<std::string, boost::shared_ptr<Joint>> joints_;
    std::string name_;
    boost::shared_ptr<Link> root_link_;
};
// urdf::Model (the XML parser) is not available: initString takes the PATH of the robot description fixture and the
// model is built programmatically by ref_collision_shim.cpp (OracleRefUrdfFromFile)
class ModelInterface;
bool OracleRefUrdfFromFile(const std::string& path, ModelInterface& model);
class Model : public ModelInterface { public: bool initString(const std::string& s) { return OracleRefUrdfFromFile(s, *this); } };
} // namespace urdf
