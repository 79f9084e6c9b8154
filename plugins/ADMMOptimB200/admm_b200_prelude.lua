-- admm_b200_prelude.lua -- run the UNCHANGED reference driver on the B200 backend:
--     ugshell -ex plugins/ADMMOptimB200/admm_b200_prelude.lua -script 3d_admm.lua -numRefs 2 ...
-- The plugin ADMMOptimB200 registers the hot-path objects of the replaced plugins under their own names
-- (DeformationEquation, Testing, VolumeDefect ...).  Objects whose names belong to ugcore are registered with the prefix
-- "B200"; this file binds the ugcore names to them WHERE THE DEFORMATION / EXTENSION SUBPROBLEM USES THEM and leaves
-- everything else (Navier-Stokes, adjoint, Drag, Sensitivity, VTK of the flow fields) on UG4's CPU objects.  The dispatch is by argument
-- type: an object made from a B200 domain stays on the GPU side.
-- [UPSTREAM-UNVERIFIED]: no Lua interpreter / ugshell exists in the build image; written against the Lua API the scripts use.

local ug_PluginRequired = PluginRequired
local replaced = { ADMMOptim = true, PLaplacian = true }      -- 3d_admm.lua:1-3 (FluidOptim stays: Navier-Stokes side)
function PluginRequired(name)
	if replaced[name] then return ug_PluginRequired("ADMMOptimB200") end
	return ug_PluginRequired(name)
end

local function is_b200(obj)
	local t = ug_class_name and ug_class_name(obj) or ""
	return string.sub(t, 1, 4) == "B200"
end

-- the deformation domain lives on the GPU as well: LoadDomain fills both, CreateRegularHierarchy refines both
local b200_domains = {}
local ug_LoadDomain = LoadDomain
function LoadDomain(dom, gridName)
	ug_LoadDomain(dom, gridName)
	local gdom = B200Domain()
	B200LoadDomain(gdom, gridName)
	b200_domains[dom] = gdom
end
local ug_CreateRegularHierarchy = util.refinement.CreateRegularHierarchy
util.refinement.CreateRegularHierarchy = function(dom, numRefs, verbose, balancerDesc)
	ug_CreateRegularHierarchy(dom, numRefs, verbose, balancerDesc)
	B200CreateRegularHierarchy(b200_domains[dom], numRefs)
end

-- bDebugOutput (3d_admm.lua:795): the UG4 grid is written by UG4, the GPU-side copy next to it (same level, current coordinates)
local ug_SaveGridLevelToFile = SaveGridLevelToFile
function SaveGridLevelToFile(grid, sh, level, filename)
	ug_SaveGridLevelToFile(grid, sh, level, filename)
	for _, gdom in pairs(b200_domains) do
		B200SaveGridLevelToFile(gdom, level, string.gsub(filename, "%.ugx$", "") .. "_b200.ugx")
	end
end

-- VTKOutput (3d_admm.lua:716): one writer object serves both sides -- print() of a B200 grid function goes to the GPU-side writer
local ug_VTKOutput = VTKOutput
function VTKOutput()
	local w = { cpu = ug_VTKOutput(), gpu = B200VTKOutput() }
	function w:clear_selection() self.cpu:clear_selection(); self.gpu:clear_selection() end
	function w:select_nodal(fcts, name) self.cpu:select_nodal(fcts, name); self.gpu:select_nodal(fcts, name) end
	function w:select_all(flag) self.cpu:select_all(flag); self.gpu:select_all(flag) end
	function w:print(filename, gf, step, time, makeConsistent)
		if is_b200(gf) then return self.gpu:print(filename, gf, step or 0, time or 0, makeConsistent or false) end
		return self.cpu:print(filename, gf, step, time, makeConsistent)
	end
	return setmetatable(w, { __index = function(t, key) return function(self, ...) return self.cpu[key](self.cpu, ...) end end })
end

-- approximation spaces: the Lagrange-1 deformation space and the piecewise-constant tensor space go to the GPU, the
-- Taylor-Hood spaces of the flow problems stay on the CPU (they are created first: add_fct decides)
local ug_ApproximationSpace = ApproximationSpace
function ApproximationSpace(dom)
	local proxy = { cpu = ug_ApproximationSpace(dom), gpu = nil, dom = dom }
	local mt = {}
	mt.__index = function(t, key)
		if key == "add_fct" then
			return function(self, fcts, ftype, order)
				local hot = (ftype == "Lagrange" and order == 1 and string.find(fcts, "u1")) or ftype == "Piecewise-Constant"
				if hot then
					self.gpu = self.gpu or B200ApproximationSpace(b200_domains[self.dom])
					if order then self.gpu:add_fct(fcts, ftype, order) else self.gpu:add_fct(fcts, ftype) end
				else
					self.cpu:add_fct(fcts, ftype, order)
				end
			end
		end
		return function(self, ...)
			local target = self.gpu or self.cpu
			return target[key](target, ...)
		end
	end
	return setmetatable(proxy, mt)
end
local function space_of(s) return (type(s) == "table" and (s.gpu or s.cpu)) or s end

local function dispatch(ug_ctor, b200_ctor)
	return function(first, ...)
		local a = space_of(first)
		if first ~= nil and is_b200(a) then return b200_ctor(a, ...) end
		return ug_ctor(a, ...)
	end
end
GridFunction = dispatch(GridFunction, B200GridFunction)
AdvancedGridFunction = dispatch(AdvancedGridFunction, B200GridFunction)
DomainDiscretization = dispatch(DomainDiscretization, B200DomainDiscretization)
AssembledLinearOperator = dispatch(AssembledLinearOperator, B200AssembledLinearOperator)
GlobalGridFunctionNumberData = dispatch(GlobalGridFunctionNumberData, B200GridFunctionNumberData)
GlobalGridFunctionGradientData = dispatch(GlobalGridFunctionGradientData, B200GridFunctionGradientData)
GeometricMultiGrid = dispatch(GeometricMultiGrid, B200GeometricMultiGrid)
VecScaleAssign = dispatch(VecScaleAssign, B200VecScaleAssign)
VecScaleAdd2 = dispatch(VecScaleAdd2, B200VecScaleAdd2)
VecProd = dispatch(VecProd, B200VecProd)
VecNorm = dispatch(VecNorm, B200VecNorm)
L2Norm = dispatch(L2Norm, B200L2Norm)

-- the solver descriptor of obstacle_optim_3d_util.lua:10-40 for deformation-space discretisations
local ug_CreateSolver = util.solver.CreateSolver
util.solver.CreateSolver = function(desc)
	local pre = desc.precond
	if type(pre) == "table" and pre.type == "gmg" and is_b200(space_of(pre.approxSpace)) then
		local gmg = B200GeometricMultiGrid(space_of(pre.approxSpace))
		gmg:set_base_level(pre.baseLevel or 0)
		gmg:set_base_solver(B200SuperLU())
		gmg:set_gathered_base_solver_if_ambiguous(pre.gatheredBaseSolverIfAmbiguous or false)
		gmg:set_smoother(B200GaussSeidel())
		gmg:set_cycle_type(pre.cycle or "V")
		gmg:set_num_presmooth(pre.preSmooth or 2)
		gmg:set_num_postsmooth(pre.postSmooth or 2)
		gmg:set_rap(pre.rap or false)
		gmg:set_discretization(pre.discretization)
		local s = B200BiCGStab()
		s:set_preconditioner(gmg)
		local cc = desc.convCheck or {}
		s:set_convergence_check(B200ConvCheck(cc.iterations or 100, cc.absolute or 1e-12, cc.reduction or 1e-12, cc.verbose or false))
		return s
	end
	return ug_CreateSolver(desc)
end
-- CG() + Jacobi(0.66) + ConvCheck(...) of 3d_admm.lua:701-703 are bound when the solver is initialised with a B200 operator:
-- the driver creates them before any operator is known, so these three constructors return thin recorders
local ug_CG, ug_Jacobi, ug_ConvCheck = CG, Jacobi, ConvCheck
function Jacobi(damp) return { damp = damp, cpu = function() return damp and ug_Jacobi(damp) or ug_Jacobi() end } end
function ConvCheck(...) local a = {...}; return { args = a, cpu = function() return ug_ConvCheck(unpack(a)) end } end
function CG()
	local rec = { pre = nil, cc = nil, impl = nil }
	function rec:set_preconditioner(p) self.pre = p end
	function rec:set_convergence_check(c) self.cc = c end
	function rec:init(A, x)
		if not self.impl then
			if is_b200(A) then
				self.impl = B200CG()
				if self.pre then self.impl:set_preconditioner(B200Jacobi(self.pre.damp or 1.0)) end
				if self.cc then self.impl:set_convergence_check(B200ConvCheck(unpack(self.cc.args))) end
			else
				self.impl = ug_CG()
				if self.pre then self.impl:set_preconditioner(self.pre.cpu()) end
				if self.cc then self.impl:set_convergence_check(self.cc.cpu()) end
			end
		end
		return self.impl:init(A, x)
	end
	function rec:apply(x, b) return self.impl:apply(x, b) end
	function rec:apply_return_defect(x, b) return self.impl:apply_return_defect(x, b) end
	function rec:step() return self.impl:step() end
	return rec
end
