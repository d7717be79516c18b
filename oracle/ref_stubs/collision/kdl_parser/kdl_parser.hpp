// ORACLE (test infrastructure only).  Stand-in for kdl_parser (absent third-party dependency): treeFromUrdfModel
// restated from its published source -- one KDL segment per URDF link below the root, named after the link, with the
// link's parent joint (origin = the joint frame's position, axis = the joint frame's rotation applied to the URDF
// axis) and the joint frame as tip; children are added in the order of the link's child_links.
#pragma once
#include <kdl/frames.hpp>
#include <urdf_model/model.h>
namespace kdl_parser {
inline KDL::Frame toKdl(const urdf::Pose& p)
{
    return KDL::Frame(KDL::Rotation::Quaternion(p.rotation.x, p.rotation.y, p.rotation.z, p.rotation.w),
                      KDL::Vector(p.position.x, p.position.y, p.position.z));
}
inline KDL::Joint toKdl(const boost::shared_ptr<urdf::Joint>& jnt)
{
    const KDL::Frame F_parent_jnt = toKdl(jnt->parent_to_joint_origin_transform);
    const KDL::Vector axis(jnt->axis.x, jnt->axis.y, jnt->axis.z);
    switch (jnt->type) {
    case urdf::Joint::FIXED: return KDL::Joint(jnt->name, KDL::Joint::None);
    case urdf::Joint::REVOLUTE:
    case urdf::Joint::CONTINUOUS: return KDL::Joint(jnt->name, F_parent_jnt.p, F_parent_jnt.M * axis, KDL::Joint::RotAxis);
    case urdf::Joint::PRISMATIC: return KDL::Joint(jnt->name, F_parent_jnt.p, F_parent_jnt.M * axis, KDL::Joint::TransAxis);
    default: return KDL::Joint(jnt->name, KDL::Joint::None);
    }
}
inline bool addChildrenToTree(const boost::shared_ptr<const urdf::Link>& root, KDL::Tree& tree)
{
    const KDL::Segment sgm(root->name, toKdl(root->parent_joint), toKdl(root->parent_joint->parent_to_joint_origin_transform));
    if (!tree.addSegment(sgm, root->parent_joint->parent_link_name)) return false;
    for (const auto& child : root->child_links) {
        if (!addChildrenToTree(child, tree)) return false;
    }
    return true;
}
inline bool treeFromUrdfModel(const urdf::ModelInterface& robot_model, KDL::Tree& tree)
{
    if (!robot_model.getRoot()) return false;
    tree = KDL::Tree(robot_model.getRoot()->name);
    for (const auto& child : robot_model.getRoot()->child_links) {
        if (!addChildrenToTree(child, tree)) return false;
    }
    return true;
}
} // namespace kdl_parser
