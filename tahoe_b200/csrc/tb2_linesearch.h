// tb2_linesearch.h -- the secant line search of the nonlinear PCG solver as a proposal / observation state machine.
//
// What it must reproduce (iteration-count parity with the reference depends on taking the same decisions): PCGSolver_LS::Update
// (PCGSolver_LS.cpp:213-348) looks for the step s along the search direction where G(s) = R(u + s dir) . dir vanishes.  It knows
// G(0), evaluates G(1), then repeats secant steps through the two bracket points; the trial replaces the bracket point of
// largest |G| when that beats both other values, else -- with both bracket values of one sign -- the point whose sign the trial
// opposes; otherwise the search stalls.  A secant root beyond max_step is clamped (one last evaluation), a negative root ends
// the search at once, and so does reaching the trial budget.  A stalled search settles on the best step tried.
//
// Host-only, no CUDA: the caller owns the evaluation of G (one element sweep per trial, tb2_nlpcg.cu).
#pragma once
#include <cmath>
#include <vector>

namespace tb2 {

class SecantSearch {
public:
    struct Trial {
        double step, slope; // s and G(s)
    };

    SecantSearch(double max_step, double abs_tolerance, double rel_tolerance, int max_trials)
        : max_step_(max_step), abs_tol_(abs_tolerance), rel_tol_(rel_tolerance), max_trials_(max_trials)
    {
        trials_.reserve(max_trials > 3 ? (size_t)max_trials + 1 : 4);
    }

    // G(0) is known from the direction update
    void begin(double slope_at_zero)
    {
        trials_.clear();
        trials_.push_back({0.0, slope_at_zero});
        lo_ = 0;
        hi_ = -1;
        phase_ = Phase::kUnitStep;
        stalled_ = false;
    }

    // the next step to evaluate; false: the search is over
    bool propose(double* step)
    {
        switch (phase_) {
        case Phase::kUnitStep:
            pending_ = 1.0;
            break;
        case Phase::kSecant: {
            const Trial &A = trials_[(size_t)lo_], &B = trials_[(size_t)hi_];
            const double rise = (A.slope - B.slope) / (A.step - B.step); // the line through the bracket, slope-intercept form
            const double at_zero = B.slope - rise * B.step;
            const double root = -at_zero / rise;
            if (root < 0.0) {
                stalled_ = true;
                phase_ = Phase::kOver;
                return false;
            }
            if (root > max_step_) { // clamp: one last evaluation at the largest admissible step
                stalled_ = true;
                phase_ = Phase::kClamped;
                pending_ = max_step_;
            } else
                pending_ = root;
            break;
        }
        default:
            return false;
        }
        *step = pending_;
        return true;
    }

    // G at the step propose() handed out
    void observe(double slope)
    {
        trials_.push_back({pending_, slope});
        const int latest = (int)trials_.size() - 1;
        if (phase_ == Phase::kUnitStep) {
            hi_ = latest;
            // the smaller of the two opening magnitudes scales the relative test
            scale_ = std::fabs(trials_[0].slope) > std::fabs(slope) ? slope : trials_[0].slope;
            phase_ = Phase::kSecant;
            return;
        }
        if (phase_ == Phase::kClamped) {
            phase_ = Phase::kOver;
            return;
        }
        const double ga = std::fabs(trials_[(size_t)lo_].slope), gb = std::fabs(trials_[(size_t)hi_].slope), gn = std::fabs(slope);
        if (ga > gn && ga > gb) lo_ = latest;
        else if (gb > gn && gb > ga) hi_ = latest;
        else if (trials_[(size_t)lo_].slope * trials_[(size_t)hi_].slope > 0) { // no sign change in the bracket yet
            if (trials_[(size_t)lo_].slope * slope < 0) lo_ = latest;
            else if (trials_[(size_t)hi_].slope * slope < 0) hi_ = latest;
            else stalled_ = true;
        } else
            stalled_ = true;
        if ((int)trials_.size() >= max_trials_) stalled_ = true;
        const bool resolved = !(gn > abs_tol_) || !(std::fabs(slope / scale_) > rel_tol_);
        if (resolved || stalled_) phase_ = Phase::kOver;
    }

    // true: the search ended without meeting its tolerances -- the caller moves to best_step()
    bool stalled() const { return stalled_; }

    // the tried step of smallest |G|, preferring any non-zero step over s = 0
    double best_step() const
    {
        size_t best = 0;
        double s_best = std::fabs(trials_[0].step), g_best = std::fabs(trials_[0].slope);
        for (size_t i = 1; i < trials_.size(); i++) {
            const double s = std::fabs(trials_[i].step), g = std::fabs(trials_[i].slope);
            if (s_best < 1.0e-12 || (s > 1.0e-12 && g < g_best)) {
                s_best = s;
                g_best = g;
                best = i;
            }
        }
        return trials_[best].step;
    }

    const std::vector<Trial>& trials() const { return trials_; }

private:
    enum class Phase { kUnitStep, kSecant, kClamped, kOver };
    double max_step_, abs_tol_, rel_tol_;
    int max_trials_;
    std::vector<Trial> trials_;
    int lo_ = 0, hi_ = -1; // the two bracket points, as indices into trials_
    double pending_ = 0.0, scale_ = 1.0;
    Phase phase_ = Phase::kOver;
    bool stalled_ = false;
};

} // namespace tb2
