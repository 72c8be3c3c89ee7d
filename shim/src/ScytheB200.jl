# ScytheB200.jl -- Julia binding of libscythe_b200.so (include/scythe_b200.h) under the reference's own names.
#
# Shadows the Springsteel surface Scythe.jl calls (SURVEY App. A.1) and the driver of src/Scythe.jl:37-62 /
# src/semiimplicit.jl:126-299.  One `ccall` per C entry point; structs mirror the header field by field
# (sb_grid_params, sb_grid_info, sb_model_params).  NOT executed in the build image (no Julia there): kept mechanical.
module ScytheB200

using Libdl

export GPUGrid, createGrid, spectralTransform!, gridTransform!, splineTransform!, tileTransform!, calcTileSizes,
       getGridpoints, num_columns, checkCFL, integrate_model, GPUModel

const libpath = get(ENV, "SCYTHE_B200_LIB", "libscythe_b200")
const lib = Ref{Ptr{Cvoid}}(C_NULL)
function __init__()
    lib[] = Libdl.dlopen(libpath)          # fails loudly when the CUDA library is missing: there is no CPU fallback
end
sym(name::Symbol) = Libdl.dlsym(lib[], name)

# ---------------------------------------------------------------------------------------------- error convention
const SB_EINVAL, SB_EDOMAIN, SB_ECUDA, SB_ENAN, SB_EUNSUPPORTED, SB_ECOMM = -1, -2, -3, -4, -5, -6
function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall(sym(:sb_last_error), Cstring, ()))
    rc == SB_EDOMAIN && throw(DomainError(0, msg))          # as src/spectralGrid.jl:88,91
    rc == SB_EINVAL && throw(ArgumentError(msg))
    error(msg)                                              # SB_ENAN: "NaN found in variable ..." (src/semiimplicit.jl:745)
end

# ---------------------------------------------------------------------------------------------- C structs
struct GridParamsC                    # sb_grid_params
    geometry::Int32
    nvars::Int32
    xmin::Float64
    xmax::Float64
    num_cells::Int64
    l_q::Float64
    zmin::Float64
    zmax::Float64
    zDim::Int64
    b_zDim::Int64
    spectralIndexL::Int64
    tile_num::Int64
    BCL::Ptr{Int32}
    BCR::Ptr{Int32}
    BCB::Ptr{Int32}
    BCT::Ptr{Int32}
end

struct GridInfoC                      # sb_grid_info
    N::Int64
    V::Int64
    D::Int64
    S::Int64
    rDim::Int64
    b_rDim::Int64
    zDim::Int64
    b_zDim::Int64
    kDim::Int64
    lDim::Int64
    num_columns::Int64
    patchOffsetL::Int64
    ndims::Int64
end

struct ModelParamsC                   # sb_model_params
    ts::Float64
    integration_time::Float64
    output_interval::Float64
    equation_set::Cstring
    grid::Ptr{GridParamsC}
    var_names::Ptr{Cstring}
    n_physical_params::Int32
    param_names::Ptr{Cstring}
    param_values::Ptr{Float64}
    semiimplicit::Int32
    ref_sbar::Ptr{Float64}
    ref_xibar::Ptr{Float64}
    ref_mubar::Ptr{Float64}
    Pxi_bar::Float64
    ref_mu_lbar::Ptr{Float64}
end

const GEOM = Dict("R" => 0, "RZ" => 1, "RL" => 2, "RLZ" => 3)

# BC dictionaries are compared by content with the reference's constants (CubicBSpline.R0 = Dict("R0" => 0), ...;
# notebooks/LinearAdvection_example.ipynb:43), so this file needs no symbol of Springsteel at load time.
function splinebc(d::AbstractDict)
    haskey(d, "PERIODIC") && return Int32(7)
    haskey(d, "R0") && return Int32(0)
    haskey(d, "R3") && return Int32(6)
    if haskey(d, "α1") || haskey(d, "alpha1")                # rank-1: R1T0 (u=0), R1T1 (u'=0), R1T2 (u''=0)
        a = get(d, "α1", get(d, "alpha1", 0.0)); b = get(d, "β1", get(d, "beta1", 0.0))
        (a, b) == (-4.0, -1.0) && return Int32(1)
        (a, b) == (0.0, 1.0) && return Int32(2)
        (a, b) == (2.0, -1.0) && return Int32(3)
    end
    if haskey(d, "α2") || haskey(d, "alpha2")                # rank-2: R2T10, R2T20
        b = get(d, "β2", get(d, "beta2", 0.0))
        return b == -1.0 ? Int32(4) : Int32(5)
    end
    for (name, code) in ("R1T0" => 1, "R1T1" => 2, "R1T2" => 3, "R2T10" => 4, "R2T20" => 5)
        haskey(d, name) && return Int32(code)
    end
    throw(ArgumentError("unknown radial boundary condition $(d)"))
end
function chebbc(d::AbstractDict)
    for (name, code) in ("R0" => 0, "R1T0" => 1, "R1T1" => 2, "R1T2" => 3)
        haskey(d, name) && return Int32(code)
    end
    if haskey(d, "α0"); return Int32(1); end
    if haskey(d, "α1"); return Int32(2); end
    if haskey(d, "α2"); return Int32(3); end
    throw(ArgumentError("unknown vertical boundary condition $(d)"))
end

"variable names ordered by column (gp.vars :: Dict{String,Int}, src/spectralGrid.jl:38)"
varnames(gp) = sort(collect(keys(gp.vars)); by = k -> gp.vars[k])

"Keeps the BC arrays alive next to the C struct that points at them."
struct GridParamsKeep
    c::GridParamsC
    bcl::Vector{Int32}
    bcr::Vector{Int32}
    bcb::Vector{Int32}
    bct::Vector{Int32}
end

function to_c(gp)::GridParamsKeep       # gp :: Springsteel.GridParameters (fields: src/spectralGrid.jl:20-45)
    names = varnames(gp)
    pick(d, k, f) = haskey(d, k) ? f(d[k]) : (haskey(d, "default") ? f(d["default"]) : Int32(0))
    bcl = Int32[pick(gp.BCL, k, splinebc) for k in names]
    bcr = Int32[pick(gp.BCR, k, splinebc) for k in names]
    hasz = gp.geometry in ("RZ", "RLZ")
    bcb = hasz ? Int32[pick(gp.BCB, k, chebbc) for k in names] : zeros(Int32, length(names))
    bct = hasz ? Int32[pick(gp.BCT, k, chebbc) for k in names] : zeros(Int32, length(names))
    haskey(GEOM, gp.geometry) || throw(DomainError(gp.geometry, "Unknown geometry"))
    c = GridParamsC(GEOM[gp.geometry], length(names), gp.xmin, gp.xmax, gp.num_cells, gp.l_q,
                    hasz ? gp.zmin : 0.0, hasz ? gp.zmax : 0.0, hasz ? gp.zDim : 0, hasz ? gp.b_zDim : 0,
                    gp.spectralIndexL, gp.tile_num, pointer(bcl), pointer(bcr), pointer(bcb), pointer(bct))
    return GridParamsKeep(c, bcl, bcr, bcb, bct)
end

# ---------------------------------------------------------------------------------------------- grids
"Drop-in for Springsteel's R_Grid / RL_Grid / RZ_Grid / RLZ_Grid: host mirrors with the reference layout."
mutable struct GPUGrid
    params::Any                           # the Springsteel GridParameters the caller passed
    handle::Ptr{Cvoid}
    info::GridInfoC
    physical::Array{Float64,3}            # [N, V, D]
    spectral::Array{Float64,2}            # [S, V]
    owned::Bool
end

function grid_info(h::Ptr{Cvoid})
    r = Ref{GridInfoC}()
    check(ccall(sym(:sb_grid_get_info), Cint, (Ptr{Cvoid}, Ref{GridInfoC}), h, r))
    return r[]
end

function createGrid(gp; device::Integer = 0)                 # src/semiimplicit.jl:130,150,155
    keep = to_c(gp)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep begin
        check(ccall(sym(:sb_grid_create), Cint, (Ref{GridParamsC}, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                    keep.c, device, C_NULL, h))
    end
    i = grid_info(h[])
    g = GPUGrid(gp, h[], i, zeros(i.N, i.V, i.D), zeros(i.S, i.V), true)
    finalizer(x -> (x.owned && x.handle != C_NULL) ? ccall(sym(:sb_grid_destroy), Cint, (Ptr{Cvoid},), x.handle) : Cint(0), g)
    return g
end

function spectralTransform!(g::GPUGrid)                      # src/semiimplicit.jl:135,734
    check(ccall(sym(:sb_grid_set_physical), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32), g.handle, g.physical, 0, 1))
    check(ccall(sym(:sb_spectral_transform), Cint, (Ptr{Cvoid},), g.handle))
    check(ccall(sym(:sb_grid_get_spectral), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), g.handle, 0, g.spectral))
    return g.spectral
end

function gridTransform!(g::GPUGrid)                          # src/semiimplicit.jl:136
    check(ccall(sym(:sb_grid_set_spectral), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), g.handle, 0, g.spectral))
    check(ccall(sym(:sb_grid_transform), Cint, (Ptr{Cvoid},), g.handle))
    check(ccall(sym(:sb_grid_get_physical), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32),
                g.handle, g.physical, 0, size(g.physical, 3)))
    return g.physical
end

# splineTransform!(patchSplines, patchSpectral, gp, sharedSpectral, tile)         src/semiimplicit.jl:237,285
# patchSplines / gp / tile carry nothing the library does not already hold for `patch`.
function splineTransform!(patch::GPUGrid, patchSpectral::Array{Float64}, sharedSpectral::AbstractArray{Float64})
    shared = sharedSpectral isa Array{Float64} ? sharedSpectral : Array{Float64}(sharedSpectral)   # SharedArray -> dense
    check(ccall(sym(:sb_grid_set_spectral), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), patch.handle, 0, shared))
    check(ccall(sym(:sb_spline_transform), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), patch.handle, patch.handle))
    check(ccall(sym(:sb_grid_get_spectral), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), patch.handle, 1, patchSpectral))
    return patchSpectral
end

# tileTransform!(patchSplines, patchSpectral, gp, tile, splineBuffer)              src/semiimplicit.jl:241,252,290,305
function tileTransform!(patch::GPUGrid, patchSpectral::Array{Float64}, tile::GPUGrid)
    check(ccall(sym(:sb_grid_set_spectral), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), patch.handle, 1, patchSpectral))
    check(ccall(sym(:sb_tile_transform), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), patch.handle, tile.handle))
    check(ccall(sym(:sb_grid_get_physical), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32),
                tile.handle, tile.physical, 0, size(tile.physical, 3)))
    return tile.physical
end

function calcTileSizes(patch::GPUGrid, n::Integer)           # src/semiimplicit.jl:141-144: rows xmin,xmax,num_cells,spectralIndexL,npts
    out = zeros(5, n)
    keep = to_c(patch.params)
    GC.@preserve keep begin
        check(ccall(sym(:sb_calc_tile_sizes), Cint, (Ref{GridParamsC}, Int32, Ptr{Float64}), keep.c, n, out))
    end
    return out
end

num_columns(g::GPUGrid) = Int(g.info.num_columns)            # src/semiimplicit.jl:308

function getGridpoints(g::GPUGrid)                           # src/semiimplicit.jl:59
    nd = Int(g.info.ndims)
    out = zeros(Int(g.info.N), nd)
    check(ccall(sym(:sb_grid_get_gridpoints), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), g.handle, out, length(out)))
    return nd == 1 ? vec(out) : out
end

function checkCFL(g::GPUGrid)                                # src/semiimplicit.jl:737-751
    v = Ref{Int32}(0); i = Ref{Int64}(0)
    check(ccall(sym(:sb_check_cfl), Cint, (Ptr{Cvoid}, Ref{Int32}, Ref{Int64}), g.handle, v, i))
end

# sharedSpectral .= 0 (:272); sharedSpectral[patchIndexMap] .= tileView, halo add (:320-329, 279-282)
shared_clear!(patch::GPUGrid) = check(ccall(sym(:sb_shared_clear), Cint, (Ptr{Cvoid},), patch.handle))
shared_assemble!(patch::GPUGrid, tile::GPUGrid, prev::Union{GPUGrid,Nothing}, last::Bool) =
    check(ccall(sym(:sb_shared_assemble), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32),
                patch.handle, tile.handle, prev === nothing ? C_NULL : prev.handle, last ? 1 : 0))

# ---------------------------------------------------------------------------------------------- model (device resident)
"initialize_model + run_model for the tiles this process owns (one tile per reference worker, src/semiimplicit.jl:155)."
mutable struct GPUModel
    handle::Ptr{Cvoid}
    patch::GPUGrid
    num_tiles::Int
    tile_first::Int
    tile_count::Int
    t::Int64
end

function GPUModel(model; num_tiles::Integer = 1, tile_first::Integer = 0, tile_count::Integer = num_tiles - tile_first,
                  device::Integer = 0, ref_state = nothing)
    gp = model.grid_params
    keep = to_c(gp)
    names = varnames(gp)
    pkeys = collect(keys(model.physical_params))
    pnames = String[String(k) for k in pkeys]                              # Dict{Symbol,Float64} keys (:g, :K, ...)
    pvals = Float64[Float64(model.physical_params[k]) for k in pkeys]
    semi = Int32(get(model.options, :semiimplicit, false) ? 1 : 0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    eq = String(model.equation_set)
    sbar = ref_state === nothing ? Float64[] : Array{Float64}(ref_state.sbar)     # [zDim, 3] value, d/dz, d2/dz2
    xibar = ref_state === nothing ? Float64[] : Array{Float64}(ref_state.xibar)
    mubar = ref_state === nothing ? Float64[] : Array{Float64}(ref_state.mubar)
    mu_lbar = ref_state === nothing ? Float64[] : Array{Float64}(ref_state.mu_lbar)   # BF02_test (src/testModels.jl:291-293)
    GC.@preserve keep names pnames pvals eq sbar xibar mubar mu_lbar begin
        cnames = Cstring[Base.unsafe_convert(Cstring, n) for n in names]
        cpn = isempty(pnames) ? Cstring[Base.unsafe_convert(Cstring, "")] : Cstring[Base.unsafe_convert(Cstring, n) for n in pnames]
        gref = Ref(keep.c)
        GC.@preserve cnames cpn gref begin
            mp = ModelParamsC(model.ts, model.integration_time, model.output_interval, Base.unsafe_convert(Cstring, eq),
                              Base.unsafe_convert(Ptr{GridParamsC}, gref), pointer(cnames), length(pnames), pointer(cpn),
                              isempty(pvals) ? C_NULL : pointer(pvals), semi,
                              isempty(sbar) ? C_NULL : pointer(sbar), isempty(xibar) ? C_NULL : pointer(xibar),
                              isempty(mubar) ? C_NULL : pointer(mubar), ref_state === nothing ? 0.0 : Float64(ref_state.Pxi_bar),
                              length(mu_lbar) == length(sbar) && !isempty(mu_lbar) ? pointer(mu_lbar) : C_NULL)
            check(ccall(sym(:sb_model_create), Cint,
                        (Ref{ModelParamsC}, Int32, Int32, Int32, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                        mp, num_tiles, tile_first, tile_count, device, C_NULL, h))
        end
    end
    ph = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall(sym(:sb_model_patch), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), h[], ph))
    i = grid_info(ph[])
    patch = GPUGrid(gp, ph[], i, zeros(i.N, i.V, i.D), zeros(i.S, i.V), false)   # borrowed handle: owned by the model
    m = GPUModel(h[], patch, num_tiles, tile_first, tile_count, 0)
    finalizer(x -> x.handle != C_NULL ? ccall(sym(:sb_model_destroy), Cint, (Ptr{Cvoid},), x.handle) : Cint(0), m)
    return m
end

initialize!(m::GPUModel, ic::Array{Float64}) =               # initialize_model, src/semiimplicit.jl:126-193; ic = [N_patch, V]
    (check(ccall(sym(:sb_model_initialize), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.handle, ic)); m.t = 0; m)
function run!(m::GPUModel, nsteps::Integer)                  # model_loop without output, :268-297
    check(ccall(sym(:sb_model_run), Cint, (Ptr{Cvoid}, Int64, Int64), m.handle, m.t + 1, nsteps))
    m.t += nsteps
    return m
end
function output!(m::GPUModel)                                # :289-291 (tileTransform!(patch) + checkCFL)
    check(ccall(sym(:sb_model_output), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.handle, m.patch.physical))
    return m.patch.physical
end
"0 = what the built-in equation-set kernel reads (default), 1 = all D slots of every variable (the reference's dataflow)"
set_k3_slots!(m::GPUModel, mode::Integer) = check(ccall(sym(:sb_model_set_k3_slots), Cint, (Ptr{Cvoid}, Int32), m.handle, mode))
get_state!(m::GPUModel, tile::Integer, which::Integer, out::Array{Float64}) =
    check(ccall(sym(:sb_model_get_state), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), m.handle, tile, which, out))
set_state!(m::GPUModel, tile::Integer, which::Integer, a::Array{Float64}) =
    check(ccall(sym(:sb_model_set_state), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), m.handle, tile, which, a))
# host-resident state between steps, pipelined (INTEGRATION.md 2a); arrays must be page-locked and live until drain
stage_in!(m::GPUModel, tile::Integer, a::Array{Float64}) =
    check(ccall(sym(:sb_model_stage_in), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), m.handle, tile, a))
stage_out!(m::GPUModel, tile::Integer, a::Array{Float64}) =
    check(ccall(sym(:sb_model_stage_out), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), m.handle, tile, a))
cycle!(m::GPUModel) = (m.t += 1; check(ccall(sym(:sb_model_cycle), Cint, (Ptr{Cvoid}, Int64), m.handle, m.t)))
drain!(m::GPUModel; block::Bool = true) = check(ccall(sym(:sb_model_stage_drain), Cint, (Ptr{Cvoid}, Int32), m.handle, block ? 1 : 0))
# multi-GPU: one Julia worker per GPU, NCCL id shipped like the RemoteChannels of src/semiimplicit.jl:205-219
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall(sym(:sb_comm_unique_id), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init!(m::GPUModel, id::Vector{UInt8}, rank::Integer, nranks::Integer) = begin
    check(ccall(sym(:sb_model_colsolve_init), Cint, (Ptr{Cvoid}, Int32, Int32), m.handle, rank, nranks))
    check(ccall(sym(:sb_model_comm_init), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), m.handle, id, rank, nranks))
end
close!(m::GPUModel) = (m.handle != C_NULL && ccall(sym(:sb_model_destroy), Cint, (Ptr{Cvoid},), m.handle); m.handle = C_NULL; nothing)

# CSV of read_physical_grid / write_grid (src/semiimplicit.jl:134, src/io.jl:3-13): coordinate columns, then one per variable
function read_physical_ic(path::AbstractString, gp, N::Integer)
    lines = readlines(path)
    header = split(strip(lines[1]), ",")
    names = varnames(gp)
    cols = [findfirst(==(n), header) for n in names]
    any(isnothing, cols) && throw(ArgumentError("initial-conditions file lacks a column for one of $(names)"))
    ic = zeros(N, length(names))
    for (i, ln) in enumerate(lines[2:end])
        f = split(ln, ",")
        for (v, c) in enumerate(cols)
            ic[i, v] = parse(Float64, f[c])
        end
    end
    return ic
end
function write_physical(path::AbstractString, pts::AbstractArray, gp, physical::Array{Float64,3})
    names = varnames(gp)
    coord = gp.geometry == "RLZ" ? ["r", "l", "z"] : gp.geometry == "RZ" ? ["r", "z"] : gp.geometry == "RL" ? ["r", "l"] : ["r"]
    P = ndims(pts) == 1 ? reshape(pts, :, 1) : pts
    open(path, "w") do io
        println(io, join(vcat(coord[1:size(P, 2)], names), ","))
        for i in 1:size(P, 1)
            println(io, join(vcat([string(P[i, j]) for j in 1:size(P, 2)], [string(physical[i, v, 1]) for v in 1:length(names)]), ","))
        end
    end
end

"src/Scythe.jl:37-62 + model_loop (src/semiimplicit.jl:258-299) on one GPU (num_tiles tiles emulate the reference's workers)."
function integrate_model(model; num_tiles::Integer = 1, device::Integer = 0, ref_state = nothing, write::Bool = true)
    num_tiles >= 1 || error("Need to add at least 1 worker process")       # src/Scythe.jl:39-41
    m = GPUModel(model; num_tiles = num_tiles, device = device, ref_state = ref_state)
    ic = read_physical_ic(model.initial_conditions, model.grid_params, Int(m.patch.info.N))
    initialize!(m, ic)
    pts = getGridpoints(m.patch)
    num_ts = round(Int, model.integration_time / model.ts)
    output_int = max(round(Int, model.output_interval / model.ts), 1)
    tag(t) = string(round(t; digits = 2))                                   # src/io.jl:5
    write && (mkpath(model.output_dir); write_physical(joinpath(model.output_dir, "physical_out_$(tag(0.0)).csv"), pts, model.grid_params, output!(m)))
    t = 0
    while t < num_ts
        n = min(output_int - t % output_int, num_ts - t)
        run!(m, n); t += n
        if t % output_int == 0                                              # :288-293
            out = output!(m)
            write && write_physical(joinpath(model.output_dir, "physical_out_$(tag(t * model.ts)).csv"), pts, model.grid_params, out)
        end
    end
    out = copy(output!(m))                                                  # finalize_model, :351-355
    write && write_physical(joinpath(model.output_dir, "physical_out_$(tag(model.integration_time)).csv"), pts, model.grid_params, out)
    close!(m)
    return out
end

end # module
