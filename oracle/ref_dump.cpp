/* oracle/ref_dump.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference (libtahoe/libtoolbox compiled by
 * oracle/build_ref.mk) in-process, the way FEExecutionManagerT::RunJob_analysis
 * does (tahoe/src/main/FEExecutionManagerT.cpp:363-612) but stepping by hand
 * (FEManagerT.h:335-371 "driven externally"), and writes the in-memory arrays
 * at full precision, because the reference's own text output carries only 12
 * digits.
 *
 *   tahoe_dump <input.xml> <outdir> [--every N] [--fint] [--lhs] [--time]
 *
 * Files written to <outdir>: raw little-endian arrays <name>.bin plus
 * manifest.txt lines "<name> <f8|i4> <dim0> <dim1>".
 *   coords, conn (group 0, 0-based), eqnos, d_<k> v_<k> a_<k> after CloseStep of step k
 *   (every N steps and the last), rhs (FormRHS at the final state, active equations),
 *   fint (--fint: InternalForceOnNode for every node, incl. prescribed dofs),
 *   lhs_r lhs_c lhs_v + msr_bindx (--lhs: tangent re-formed at the final state,
 *   MSRMatrixT-derived matrices only), j2_alloc j2_flags j2_data (if any element is allocated; taken right
 *   after the last CloseStep, before the extra evaluations above), iters (SolverT::IterationNumber per step), iters_ic
 *   (the same after FEManagerT::InitialCondition),
 *   timing (--time: wall seconds of the step loop, steps, elements)
 *   --contact: for the first contact_3D_penalty group, at every dumped step k: cpairs_<k> (facet nodes 1-3 + striker, 0-based),
 *   carea_<k> (the pair's striker area), crhs_<k> (that group's FormRHS alone, active equations) and once cparams
 *   (penalty stiffness, friction coefficient, friction epsilon, viscous damping), cfacets / cfacet_surface (the triangulated contact
 *   surfaces), cstrikers / cstriker_area (striker nodes and their tributary areas)
 */
#include <sys/stat.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "CommunicatorT.h"
#include "ElementBaseT.h"
#include "ElementCardT.h"
#include "FEManagerT.h"
#include "FieldT.h"
#include "MSRMatrixT.h"
#include "NodeManagerT.h"
#include "ParameterListT.h"
#include "PenaltyContact3DT.h"
#include "RaggedArray2DT.h"
#include "SolverT.h"
#include "TimeManagerT.h"
#include "ofstreamT.h"

using namespace Tahoe;

static std::string g_out;
static FILE* g_manifest = NULL;

static void put(const char* name, const char* type, const void* p, size_t elsize, long d0, long d1)
{
    std::string path = g_out + "/" + name + ".bin";
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror(path.c_str()); exit(2); }
    fwrite(p, elsize, (size_t)(d0 * (d1 ? d1 : 1)), f);
    fclose(f);
    fprintf(g_manifest, "%s %s %ld %ld\n", name, type, d0, d1);
    fflush(g_manifest);
}
static void put2(const std::string& name, const dArray2DT& a) { put(name.c_str(), "f8", a.Pointer(), 8, a.MajorDim(), a.MinorDim()); }
static void put2(const std::string& name, const iArray2DT& a) { put(name.c_str(), "i4", a.Pointer(), 4, a.MajorDim(), a.MinorDim()); }
static void put1(const std::string& name, const dArrayT& a) { put(name.c_str(), "f8", a.Pointer(), 8, a.Length(), 0); }
static void put1(const std::string& name, const iArrayT& a) { put(name.c_str(), "i4", a.Pointer(), 4, a.Length(), 0); }

/* exposes the protected MSR arrays (MSRMatrixT.h:117-129) */
struct MSRPeek : public MSRMatrixT {
    static void dump(const MSRMatrixT& m)
    {
        const MSRPeek& p = static_cast<const MSRPeek&>(m);
        iArrayT r, c;
        dArrayT v;
        p.GenerateRCV(r, c, v, -1.0);
        put1("lhs_r", r);
        put1("lhs_c", c);
        put1("lhs_v", v);
        put1("msr_bindx", p.fbindx);
    }
};

/* exposes the protected contact data (ContactT.h:145-156, PenaltyContact3DT.h) */
struct ContactPeek : public PenaltyContact3DT {
    static void dump(PenaltyContact3DT& g, FEManagerT* tahoe, int solver_group, int k)
    {
        ContactPeek& p = static_cast<ContactPeek&>(g);
        char buf[64];
        const iArray2DT& pairs = *p.fConnectivities[0]; /* the active striker-facet pairs (Contact3DT::SetActiveInteractions, Contact3DT.cpp:100-175) */
        snprintf(buf, sizeof buf, "cpairs_%d", k);
        put(buf, "i4", pairs.Pointer(), 4, pairs.MajorDim(), pairs.MinorDim());
        std::vector<double> area((size_t)pairs.MajorDim());
        for (int i = 0; i < pairs.MajorDim(); i++) area[i] = p.fStrikerArea[p.fStrikerTags_map.Map(pairs(i, pairs.MinorDim() - 1))];
        snprintf(buf, sizeof buf, "carea_%d", k);
        put(buf, "f8", area.data(), 8, (long)area.size(), 0);
        double prm[4] = {p.fK, p.fMu, p.fFrictionEps, p.fViscousDamping};
        put("cparams", "f8", prm, 8, 4, 0);
        /* what the search works on (ContactT.h:145-156): the triangulated surfaces, the striker nodes, their areas */
        {
            std::vector<int> facets, surf;
            for (int sfc = 0; sfc < p.fSurfaces.Length(); sfc++)
                for (int f = 0; f < p.fSurfaces[sfc].MajorDim(); f++) {
                    for (int a = 0; a < 3; a++) facets.push_back(p.fSurfaces[sfc](f, a));
                    surf.push_back(sfc);
                }
            put("cfacets", "i4", facets.data(), 4, (long)surf.size(), 3);
            put("cfacet_surface", "i4", surf.data(), 4, (long)surf.size(), 0);
            put1("cstrikers", p.fStrikerTags);
            put1("cstriker_area", p.fStrikerArea);
        }
        /* the group's own residual contribution at the current state */
        SolverT* solver = tahoe->Solver(solver_group);
        dArrayT& rhs = const_cast<dArrayT&>(tahoe->RHS(solver_group));
        dArrayT keep(rhs);
        rhs = 0.0;
        solver->UnlockRHS();
        g.FormRHS();
        solver->LockRHS();
        snprintf(buf, sizeof buf, "crhs_%d", k);
        put1(buf, rhs);
        rhs = keep;
    }
};

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: tahoe_dump input.xml outdir [--every N] [--fint] [--lhs] [--time]\n"); return 2; }
    StringT input_file(argv[1]);
    g_out = argv[2];
    int every = 0;
    bool want_fint = false, want_lhs = false, want_time = false, want_contact = false;
    for (int i = 3; i < argc; i++) {
        if (!strcmp(argv[i], "--every") && i + 1 < argc) every = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fint")) want_fint = true;
        else if (!strcmp(argv[i], "--lhs")) want_lhs = true;
        else if (!strcmp(argv[i], "--time")) want_time = true;
        else if (!strcmp(argv[i], "--contact")) want_contact = true;
    }
    mkdir(g_out.c_str(), 0755);
    g_manifest = fopen((g_out + "/manifest.txt").c_str(), "w");

    CommunicatorT::SetArgv(&argc, &argv);
    CommunicatorT comm;
    ArrayT<StringT> options(0);

    StringT outfilename;
    outfilename.Root(input_file);
    outfilename.Append(".out");
    ofstreamT out;
    out.open(outfilename);

    int status = 0;
    try {
        ParameterListT valid_list;
        FEManagerT::ParseInput(input_file, valid_list, true, false, false, options);
        FEManagerT* tahoe = FEManagerT::New(valid_list.Name(), input_file, out, comm, options, FEManagerT::kRun);
        tahoe->TakeParameterList(valid_list);

        NodeManagerT* nodes = tahoe->NodeManager();
        FieldT* field = nodes->Field("displacement");
        ElementBaseT* group = tahoe->ElementGroup(0);
        const int solver_group = 0;

        put2("coords", nodes->InitialCoordinates());
        AutoArrayT<const iArray2DT*> c1;
        AutoArrayT<const RaggedArray2DT<int>*> c2;
        group->ConnectsU(c1, c2);
        {
            /* blocks concatenated in block order = element order (ElementBaseT.cpp:607-632) */
            std::vector<int> all;
            int nen = 0;
            for (int b = 0; b < c1.Length(); b++) {
                nen = c1[b]->MinorDim();
                all.insert(all.end(), c1[b]->Pointer(), c1[b]->Pointer() + c1[b]->Length());
            }
            put("conn", "i4", all.data(), 4, nen ? (long)all.size() / nen : 0, nen);
        }
        put2("eqnos", field->Equations());
        PenaltyContact3DT* contact = NULL;
        if (want_contact)
            for (int g = 0; g < tahoe->NumElementGroups() && !contact; g++) contact = dynamic_cast<PenaltyContact3DT*>(tahoe->ElementGroup(g));

        auto dump_fields = [&](int k) {
            char buf[64];
            for (int o = 0; o <= field->Order(); o++) {
                snprintf(buf, sizeof buf, "%c_%d", "dva"[o], k);
                put2(buf, (*field)[o]);
            }
        };

        /* the reference's own step loop (FEManagerT::Solve, FEManagerT.cpp:138-268) */
        ExceptionT::CodeT error = tahoe->InitialCondition();
        dump_fields(0);
        {   /* iterations of the solve FEManagerT::InitialCondition itself runs when the load is on at t = 0 (e.g. beam.PCG.xml) */
            int ic_iters = tahoe->Solver(solver_group)->IterationNumber();
            put("iters_ic", "i4", &ic_iters, 4, 1, 0);
        }
        TimeManagerT* tm = tahoe->TimeManager();
        auto t0 = std::chrono::steady_clock::now();
        int nsteps = 0;
        std::vector<int> iters;
        while (error == ExceptionT::kNoError && tm->Step()) {
            error = tahoe->InitStep();
            if (error == ExceptionT::kNoError) error = tahoe->SolveStep();
            if (error == ExceptionT::kNoError) error = tahoe->CloseStep();
            nsteps++;
            iters.push_back(tahoe->Solver(solver_group)->IterationNumber());
            if (error != ExceptionT::kNoError) break;
            int k = tm->StepNumber();
            if ((every > 0 && k % every == 0) || k == tm->NumberOfSteps()) {
                dump_fields(k);
                if (contact) ContactPeek::dump(*contact, tahoe, solver_group, k);
            }
        }
        auto t1 = std::chrono::steady_clock::now();
        if (error != ExceptionT::kNoError) {
            fprintf(stderr, "tahoe_dump: step loop ended on exception %d\n", (int)error);
            status = 1;
        }
        /* J2 history as stored in ElementCardT (J2SimoC0HardeningT.cpp:312-333,429-452) */
        {
            int ne = group->NumElements(), nalloc = 0, dlen = 0, ilen = 0;
            for (int e = 0; e < ne; e++)
                if (group->ElementCard(e).IsAllocated()) {
                    nalloc++;
                    dlen = group->ElementCard(e).DoubleData().Length();
                    ilen = group->ElementCard(e).IntegerData().Length();
                }
            if (nalloc > 0) {
                std::vector<int> alloc(ne, 0), flags((size_t)ne * ilen, 0);
                std::vector<double> data((size_t)ne * dlen, 0.0);
                for (int e = 0; e < ne; e++) {
                    const ElementCardT& card = group->ElementCard(e);
                    if (!card.IsAllocated()) continue;
                    alloc[e] = 1;
                    memcpy(&flags[(size_t)e * ilen], card.IntegerData().Pointer(), ilen * sizeof(int));
                    memcpy(&data[(size_t)e * dlen], card.DoubleData().Pointer(), dlen * sizeof(double));
                }
                put("j2_alloc", "i4", alloc.data(), 4, ne, 0);
                put("j2_flags", "i4", flags.data(), 4, ne, ilen);
                put("j2_data", "f8", data.data(), 8, ne, dlen);
            }
        }
        put("iters", "i4", iters.data(), 4, (long)iters.size(), 0);
        if (want_time) {
            double tv[3] = {std::chrono::duration<double>(t1 - t0).count(), (double)nsteps, (double)group->NumElements()};
            put("timing", "f8", tv, 8, 3, 0);
        }

        /* residual at the final state: what SolverT sees after FormRHS (NLSolver.cpp:75-85) */
        if (status == 0) {
            SolverT* solver = tahoe->Solver(solver_group);
            dArrayT& rhs = const_cast<dArrayT&>(tahoe->RHS(solver_group));
            rhs = 0.0;
            solver->UnlockRHS();
            tahoe->FormRHS(solver_group);
            solver->LockRHS();
            put1("rhs", rhs);
        }
        /* nodal internal force incl. prescribed dofs: SolidElementT::AddNodalForce (SolidElementT.cpp:107-200)
         * calls FormKd(+constKd), i.e. it returns +fint (and +M a when the integrator forms Ma) */
        if (status == 0 && want_fint) {
            int nn = nodes->NumNodes();
            dArray2DT fint(nn, field->NumDOF());
            dArrayT f(field->NumDOF());
            for (int n = 0; n < nn; n++) {
                tahoe->InternalForceOnNode(*field, n, f);
                for (int i = 0; i < f.Length(); i++) fint(n, i) = f[i];
            }
            put2("fint", fint);
        }
        if (status == 0 && want_lhs) {
            SolverT* solver = tahoe->Solver(solver_group);
            GlobalMatrixT& lhs = const_cast<GlobalMatrixT&>(tahoe->LHS(solver_group));
            MSRMatrixT* msr = dynamic_cast<MSRMatrixT*>(&lhs);
            if (msr) {
                lhs.Clear();
                solver->UnlockLHS();
                tahoe->FormLHS(solver_group, tahoe->GlobalSystemType(solver_group));
                solver->LockLHS();
                MSRPeek::dump(*msr);
            } else
                fprintf(stderr, "tahoe_dump: --lhs needs an MSRMatrixT-derived matrix (e.g. SPOOLES_matrix)\n");
        }
        delete tahoe;
    } catch (ExceptionT::CodeT code) {
        fprintf(stderr, "tahoe_dump: exception %d\n", (int)code);
        status = 1;
    }
    fclose(g_manifest);
    return status;
}
