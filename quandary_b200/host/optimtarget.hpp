/* optimtarget.hpp -- initial conditions, target state and final-time objective; mirrors the reference's OptimTarget
 * (include/optimtarget.hpp, src/optimtarget.cpp:1-897).  O(N^2) host work per initial condition. */
#pragma once
#include "gate.hpp"
#include "mastereq.hpp"

class OptimTarget {
 public:
  OptimTarget(std::vector<std::string> target_str, const std::string& objective_str, std::vector<std::string> initcond_str, MasterEq* mastereq,
              double total_time, std::vector<double> read_gate_rot, Vec& rho_t0, bool quietmode_);
  ~OptimTarget();

  TargetType getTargetType() const { return target_type; }
  ObjectiveType getObjectiveType() const { return objective_type; }
  /* = prepareInitialState (src/optimtarget.cpp:450-698): fills rho0 for initial condition iinit, returns its identifier */
  int prepareInitialState(int iinit, int ninit, const std::vector<size_t>& nlevels, const std::vector<size_t>& nessential, Vec& rho0);
  /* = prepareTargetState (:701-709): V rho(0) V^dagger for gate targets + purity of rho(0) */
  void prepareTargetState(const Vec& rho_t0);
  void evalJ(const Vec& state, double* J_re_ptr, double* J_im_ptr);                                   /* :713-797 */
  void evalJ_diff(const Vec& state, Vec& statebar, double J_re_bar, double J_im_bar);                  /* :800-862 */
  double finalizeJ(double obj_cost_re, double obj_cost_im);                                            /* :864-880 */
  void finalizeJ_diff(double obj_cost_re, double obj_cost_im, double* obj_cost_re_bar, double* obj_cost_im_bar);   /* :883-897 */
  /* finalizeJ(evalJ(x)) of the CURRENT target (after prepareTargetState) as c0 + <a, x> + <d, x.^2>: the form the engine's
   * penalty integral takes (TimeStepper::penaltyIntegral, src/timestepper.cpp:256-300).  a, d: [2*dim]. */
  void penaltyFunctional(Vec& a, double& c0, Vec& d) const;
  /* Schroedinger + Jtrace: J is quadratic in the overlap, 1 - (Re<tau,x>/purity)^2 - (Im<tau,x>)^2 (finalizeJ, :864-880); penaltyFunctional
   * then returns a = tau and the engine takes qb_penalty_desc.overlap = 1 */
  bool penaltyIsOverlap() const { return objective_type == ObjectiveType::JTRACE && lindbladtype == LindbladType::NONE; }
  void HilbertSchmidtOverlap(const Vec& state, bool scalebypurity, double* HS_re_ptr, double* HS_im_ptr);          /* :343-402 */
  void HilbertSchmidtOverlap_diff(Vec& statebar, bool scalebypurity, double HS_re_bar, double HS_im_bar);          /* :404-447 */
  /* what the engine's device-side objective needs after prepareTargetState (qb_engine_set_objective) */
  const Vec& getTargetState() const { return targetstate; }
  double getPurity() const { return purity_rho0; }
  long getPureStateID() const { return target_type == TargetType::PURE ? purestateID : -1; }

 private:
  long dim, dim_rho, dim_ess;
  size_t noscillators;
  TargetType target_type;
  ObjectiveType objective_type;
  InitialConditionType initcond_type;
  std::vector<size_t> initcond_IDs;
  LindbladType lindbladtype;
  Gate* targetgate = nullptr;
  double purity_rho0 = 1.0;
  long purestateID = -1;
  Vec targetstate;
  bool quietmode;
};
