// CPU-only compile test (tests/test_cv_adapter_compiles.py): include/emf_b200_cv.hpp next to the REFERENCE's own class headers
// (include/EMFusion/core/TSDF.h, ObjTSDF.h from /root/reference, against the type stand-in oracle/shim_full): every public
// method of emf::TSDF / emf::ObjTSDF that src/core/EMFusion.cpp calls must exist in the adapter with the same signature --
// argument types, constness, return type -- checked one by one with static_asserts on the member-function types.
#include "EMFusion/core/TSDF.h"
#include "EMFusion/core/ObjTSDF.h"

#define EMF_B200_CV_NAMESPACE emf_b200_cv
#include "emf_b200_cv.hpp"

#include <list>
#include <type_traits>
#include <utility>

template <typename T> struct sig;
template <typename C, typename R, typename... A> struct sig<R (C::*)(A...)> { using type = R(A...); static constexpr bool is_const = false; };
template <typename C, typename R, typename... A> struct sig<R (C::*)(A...) const> { using type = R(A...); static constexpr bool is_const = true; };

#define SAME(cls, m)                                                                                                         \
    static_assert(std::is_same<sig<decltype(&emf::cls::m)>::type, sig<decltype(&emf_b200_cv::cls::m)>::type>::value,        \
                  #cls "::" #m ": signature differs from the reference's");                                                \
    static_assert(sig<decltype(&emf::cls::m)>::is_const == sig<decltype(&emf_b200_cv::cls::m)>::is_const, #cls "::" #m ": constness differs")

// emf::TSDF (include/EMFusion/core/TSDF.h:50-268)
SAME(TSDF, reset); SAME(TSDF, getCorners); SAME(TSDF, getVolumeSize); SAME(TSDF, getVolumeRes); SAME(TSDF, getVoxelSize);
SAME(TSDF, getTruncDist); SAME(TSDF, getPose); SAME(TSDF, integrate); SAME(TSDF, updateGradients); SAME(TSDF, raycast);
SAME(TSDF, computeAssociation); SAME(TSDF, prepareTracking); SAME(TSDF, computeGradients); SAME(TSDF, computeTSDFVals);
SAME(TSDF, computeTSDFWeights); SAME(TSDF, computeHuberWeights); SAME(TSDF, normalizeTSDFWeights); SAME(TSDF, combineWeights);
SAME(TSDF, computeHessians); SAME(TSDF, reduceHessians); SAME(TSDF, computePoseUpdate); SAME(TSDF, syncTrack);
SAME(TSDF, getHuberWeights); SAME(TSDF, getTrackingWeights); SAME(TSDF, getMesh); SAME(TSDF, getTSDF); SAME(TSDF, getWeightsVol);
// emf::ObjTSDF (include/EMFusion/core/ObjTSDF.h:44-195)
SAME(ObjTSDF, getID); SAME(ObjTSDF, reset); SAME(ObjTSDF, getExProb);
SAME(ObjTSDF, updateExProb); SAME(ObjTSDF, updateClassProbs); SAME(ObjTSDF, resize); SAME(ObjTSDF, integrateMask);
SAME(ObjTSDF, computeAssociation); SAME(ObjTSDF, raycast); SAME(ObjTSDF, syncTrack); SAME(ObjTSDF, getFgProbVals);
SAME(ObjTSDF, getClassID); SAME(ObjTSDF, getMesh); SAME(ObjTSDF, getFgProbVol); SAME(ObjTSDF, getFgVolMask);

// (operator== / != take the class itself: compared by use)
static_assert(std::is_same<decltype(std::declval<const emf_b200_cv::ObjTSDF&>() == std::declval<const emf_b200_cv::ObjTSDF&>()), bool>::value, "operator==");
static_assert(std::is_same<decltype(std::declval<const emf_b200_cv::ObjTSDF&>() != std::declval<const emf_b200_cv::ObjTSDF&>()), bool>::value, "operator!=");
// constructors, inheritance, which methods are virtual
static_assert(std::is_constructible<emf_b200_cv::TSDF, cv::Vec3i, float, float, cv::Affine3f, emf::TSDFParams, cv::Size>::value, "TSDF ctor");
static_assert(std::is_constructible<emf_b200_cv::ObjTSDF, cv::Vec3i, float, float, cv::Affine3f, emf::TSDFParams, cv::Size>::value, "ObjTSDF ctor");
static_assert(std::is_base_of<emf_b200_cv::TSDF, emf_b200_cv::ObjTSDF>::value, "ObjTSDF derives from TSDF");
static_assert(std::is_copy_constructible<emf_b200_cv::ObjTSDF>::value, "ObjTSDF is copied by value into EMFusion::objects (src/core/EMFusion.cpp:549)");
static_assert(std::is_polymorphic<emf_b200_cv::TSDF>::value, "reset / getMesh / getTSDF / getWeightsVol are virtual");

// the way src/core/EMFusion.cpp uses them (instantiates the inline bodies: everything must compile, nothing is run)
void use(emf_b200_cv::TSDF& bg, std::list<emf_b200_cv::ObjTSDF>& objects, cv::cuda::GpuMat& img, cv::cuda::GpuMat& img3, cv::cuda::GpuMat& m8,
         cv::Affine3f& pose, const cv::Matx33f& intr, cv::cuda::Stream& st) {
    bg.computeAssociation(img3, pose, img, st);
    bg.prepareTracking(pose, st);
    bg.computeGradients(img3); bg.computeTSDFVals(img3); bg.computeTSDFWeights(img3); bg.computeHuberWeights(); bg.normalizeTSDFWeights();
    bg.combineWeights(img); bg.computeHessians(); bg.reduceHessians(); bg.computePoseUpdate(img3);
    bg.syncTrack(pose);
    bg.raycast(pose, intr, img, img3, img3, m8, st);
    bg.integrate(img, img, pose, intr, st);
    bg.updateGradients(st);
    for (auto& o : objects) {
        o.computeAssociation(img3, pose, img, st);
        o.raycast(pose, intr, img, img3, img3, m8, st);
        o.integrateMask(m8, m8, pose, intr, st);
        o.syncTrack(pose);
        (void)o.resize(cv::Vec3f(0, 0, 0), cv::Vec3f(1, 1, 1), 2.0f);
        (void)o.getID(); (void)o.getClassID(); (void)o.getExProb();
    }
    emf_b200_cv::ObjTSDF copy = objects.front();          // shallow copy shares the device storage
    objects.push_back(copy);
    (void)bg.getTSDF(); (void)bg.getMesh();
}
