# SequencerJB200.jl -- Julia host side of the B200-native Sequencer hot path.
#
# Drop-in for the data-parallel core of SequencerJ.jl: same exported names and keyword arguments
# (`sequence(A; scales, metrics, grid)`, `SequencerResult`, `D`, `mst`, `elong`, `order`, `L2`,
# `WASS1D`, `KLD`, `ENERGY`, `ALL_METRICS`, `elongation`, `prettyp`, `emd`, `energy`, `ensuregrid!`; the metric
# constants are the same Distances.jl / EMD / Energy objects), with `_sequence` (reference src/sequencer.jl:184-276)
# replaced by ONE `ccall` into libseqj_b200.so (C ABI in include/seqj.h).  Julia is not installed in the
# build image, so this file is syntax-reviewed only; the identical call sequence is exercised through
# the Python ctypes mirror (sequencerj.jl_b200/api.py) by tests/test_gpu_*.py.
#
# Usage:  ENV["SEQJ_B200_LIB"] = "/path/to/libseqj_b200.so";  using SequencerJB200;  r = sequence(A; scales=(1,2,4))
module SequencerJB200

using Printf: @sprintf
using Logging
using LightGraphs
using SimpleWeightedGraphs
using SparseArrays: sparse
using Distances                # the reference's metric constants ARE Distances.jl instances (src/distancemetrics.jl:234,239)

# the reference's export list (src/SequencerJ.jl:24-31; `loss` is exported there but defined nowhere in src/)
export sequence, autoscale, D, mst, elong, elongation, order, prettyp, ensuregrid!
export emd, energy
export SequencerResult, EMD, Energy
export L2, WASS1D, KLD, ENERGY, ALL_METRICS

const libseqj = get(ENV, "SEQJ_B200_LIB", joinpath(@__DIR__, "..", "libseqj_b200.so"))

"Smallest allowed graph weight (reference src/distancemetrics.jl:3)."
const ϵ = 1e-6
"Euclidean recompute threshold (reference src/distancemetrics.jl:6)."
const L2THR = 1e-7

# Metric instances are the reference's own (src/distancemetrics.jl:234-257): Distances.jl objects for Euclidean / KL,
# the package's EMD / Energy types for the CDF distances.  `metric_id` maps them onto include/seqj.h (SEQJ_L2 ..).
"""
    EMD, Energy

The reference's metric types (src/distancemetrics.jl:16-21, 121-132).  Passed as TYPES in `metrics`
(`metrics = (EMD, Energy)`, with or without `grid`) they select the grid-aware branch of the reference
(src/sequencer.jl:217-219): metric ids 4 / 5 + seqj_set_grid.  Their INSTANCES `EMD()` / `Energy()` are what WASS1D /
ENERGY are: the reference never hands them the grid (quirk Q1), each chunk gets its own unit grid.
"""
struct EMD <: SemiMetric
    grids::Union{Nothing,Tuple}
end
EMD() = EMD(nothing)
struct Energy <: SemiMetric
    grids::Union{Nothing,Tuple}
end
Energy() = Energy(nothing)

const L2 = Distances.Euclidean(L2THR)              # reference src/distancemetrics.jl:234
const KLD = Distances.KLDivergence()               # :239
const WASS1D = EMD()                               # :247
const ENERGY = Energy()                            # :254
const ALL_METRICS = (L2, WASS1D, KLD, ENERGY)      # :257
"Less verbose output for PeriodicEuclidean (reference src/distancemetrics.jl:260)."
Base.show(io::IO, s::PeriodicEuclidean) = write(io, "PeriodicEuclidean($(length(s.periods))) [$(minimum(s.periods)),$(maximum(s.periods))]")

metric_id(::Distances.Euclidean) = Int32(0)        # any threshold: `Euclidean(thr)` passes thr as l2_thresh
metric_id(::Distances.KLDivergence) = Int32(2)
metric_id(::Distances.PeriodicEuclidean) = Int32(6)   # periods are handed to the library by `_sequence`
metric_id(::Distances.Cityblock) = Int32(7)           # further element-wise Distances.jl metrics (direct tile kernel)
metric_id(::Distances.TotalVariation) = Int32(8)
metric_id(::Distances.Chebyshev) = Int32(9)
metric_id(::Distances.SqEuclidean) = Int32(10)
metric_id(::Type{EMD}) = Int32(4)
metric_id(::Type{Energy}) = Int32(5)
metric_id(m::EMD) = m.grids === nothing ? Int32(1) : error("EMD with explicit grids is a pair functor; pass the TYPE `EMD` and `grid=` to sequence")
metric_id(m::Energy) = m.grids === nothing ? Int32(3) : error("Energy with explicit grids is a pair functor; pass the TYPE `Energy` and `grid=` to sequence")
metric_id(m) = error("unsupported metric $(m): the B200 path implements Euclidean, KLDivergence, PeriodicEuclidean, Cityblock, TotalVariation, Chebyshev, SqEuclidean, EMD and Energy (no CPU fallback)")
l2_threshold(metrics) = (t = L2THR; for m in metrics; m isa Distances.Euclidean && (t = m.thresh); end; t)

"Scales used in the autoscale option (reference src/sequencer.jl:148)."
const FIB = [1,2,3,5,8,13,21,34,55,89,144,233,377,610,987,1597,2584,5168]

struct SequencerResult                                # reference src/sequencer.jl:13-20
    EOSeg::Dict{Tuple,Any}
    EOAlgScale::Dict{Tuple,Any}
    D::AbstractMatrix
    mst::LightGraphs.AbstractGraph
    η::Real
    order::AbstractVector
end
D(r::SequencerResult) = r.D
mst(r::SequencerResult) = r.mst
elong(r::SequencerResult) = r.η
order(r::SequencerResult) = r.order
Base.show(io::IO, s::SequencerResult) =
    write(io, "Sequencer Result: η = $(@sprintf("%.4g", elong(s))), order = $(prettyp(order(s),5))")

function prettyp(v::AbstractVector, len::Int = 3)    # reference src/utils.jl:48-53
    length(v) < 7 && return join(string.(v), ",")
    head = join(string.(v[1:len]), ",")
    tail = join(string.(v[(end-(len-1)):end]), ",")
    return join([head, tail], "...")
end

"Walk the graph from `idx`, returning the visited vertices (reference src/utils.jl:15-24)."
function unroll(G::AbstractGraph, idx::Int; visited = Int[])
    idx in visited && return visited
    push!(visited, idx)
    N = outneighbors(G, idx)
    length(N) < 1 && length(visited) == nv(G) ? (return visited) : nothing
    for n in N
        unroll(G, n, visited = visited)
    end
    return reverse(visited)
end

"Print the sequence of outbound nodes of a graph starting from the given index (reference src/utils.jl:5-9)."
function prettyp(G::AbstractGraph, strtidx::Int; pystyle = false)
    arrow = G isa LightGraphs.SimpleDiGraph ? "->" : "--"
    v = unroll(G, strtidx)
    println(join(string.(pystyle ? reverse(v .- 1) : v), arrow))
end

"""
    elongation(g, startidx)

Half-length over half-width of the tree `g` seen from `startidx` (reference src/sequencer.jl:348-354): hop distances
with unit weights, `hlen = mean(h)`, `hwid = mean(count per level) / 2`, both clamped to `[eps(), floatmax(Float32)]`.
A host-side helper on an N-vertex graph (the hot path computes the same integers on the device, K6).
"""
function elongation(G, startidx)
    h = gdistances(G, startidx)                      # = dijkstra_shortest_paths(G, startidx, DefaultDistance()).dists
    hlen = clamp(sum(h) / length(h), eps(), floatmax(Float32))
    levels = unique(h)
    hwid = clamp(sum(count(==(k), h) for k in levels) / length(levels) / 2, eps(), floatmax(Float32))
    return hlen / hwid + 1
end

function ensuregrid!(A, grid = nothing)::AbstractVector     # reference src/sequencer.jl:296-304
    if isnothing(grid)
        grid = float(collect(axes(A, 1)))
    end
    @assert length(grid) == size(A, 1) "Grid length must match dims 1 (row count) of A. Got grid:$(length(grid)) and A:$(size(A, 1))"
    grid
end

# ---- thin ccall layer -------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    # devices: nothing (current device) or a vector of CUDA device ordinals; with several, this one process
    # drives all of them (groups sharded over host threads inside the library, fuse on devices[1])
    function Handle(devices::Union{Nothing,AbstractVector{<:Integer}} = nothing)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        devs = devices === nothing ? Cint[] : Cint[d for d in devices]
        st = isempty(devs) ?
            ccall((:seqj_create, libseqj), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), ref, C_NULL, 0) :
            ccall((:seqj_create, libseqj), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), ref, devs, length(devs))
        h = new(ref[])
        st == 0 || error("seqj_create failed ($st): " * lasterror(h))   # no CPU fallback
        finalizer(x -> ccall((:seqj_destroy, libseqj), Cvoid, (Ptr{Cvoid},), x.ptr), h)
        h
    end
end
lasterror(h::Handle) = unsafe_string(ccall((:seqj_last_error, libseqj), Cstring, (Ptr{Cvoid},), h.ptr))
const _handle = Ref{Union{Nothing,Handle}}(nothing)
handle() = (_handle[] === nothing && (_handle[] = Handle()); _handle[])

function check(h::Handle, st::Integer)
    st == 0 && return
    msg = lasterror(h)
    st == -6 && throw(AssertionError(msg))            # SEQJ_E_DOMAIN: reference src/sequencer.jl:111
    st == -1 && throw(AssertionError(msg))            # bad scale etc.: reference src/sequencer.jl:385
    error("seqj error $st: $msg")
end

bfs_digraph(parents::Vector{Int32}, root::Integer) = begin   # tree(parents) of LightGraphs.bfs_tree
    g = SimpleDiGraph(length(parents))
    for v in eachindex(parents)
        v == root + 1 && continue
        add_edge!(g, parents[v] + 1, v)
    end
    g
end

"""
    sequence(A; scales=nothing, metrics=ALL_METRICS, grid=nothing)

Same contract as the reference's `sequence` (src/sequencer.jl:105-145); the whole of `_sequence`
runs on the GPU behind one `ccall`.
"""
function sequence(A::VecOrMat{T}; scales = nothing, metrics = ALL_METRICS, grid = nothing) where {T <: Real}
    A = A isa Vector ? hcat(A...) : A
    @info "Sequencing data with\n    shape: $(size(A))\n    metric(s): $(metrics)\n    scale(s): $(scales)"
    tuplify(x::PreMetric) = tuple(x)                  # reference src/sequencer.jl:126-127
    tuplify(x::Type) = tuple(x)
    tuplify(x) = tuple(x...)
    G = ensuregrid!(A, grid)
    metrics = (metrics !== nothing) ? tuplify(metrics) : ALL_METRICS
    A64 = A isa Matrix{Float64} ? A : convert(Matrix{Float64}, A)    # Q2: FP64 throughout
    if scales === nothing
        with_logger(NullLogger()) do
            scales = autoscale(A64, metrics = metrics, grid = G)
        end
        @info "After autoscaling, best scale = $(scales)"
    end
    scales = tuplify(scales)
    r, inmin = _sequence(A64, scales, metrics, G)
    # the reference clamps the caller's array in place (src/sequencer.jl:130); the device already
    # computed on clamped values, so touch host memory only when something was below eps()
    # (any float eltype, as `clamp!(A, eps(), typemax(eltype(A)))` does; integer matrices make the reference throw)
    inmin < eps() && eltype(A) <: AbstractFloat && clamp!(A, eltype(A)(eps()), typemax(eltype(A)))
    return r
end

function _sequence(A::Matrix{Float64}, scales, metrics, grid)
    h = handle()
    M, N = size(A)
    mids = Int32[metric_id(m) for m in metrics]
    scs = Int32[s for s in scales]
    if any(id -> id == 4 || id == 5, mids)                 # the TYPES EMD / Energy: distances on `grid`
        g64 = Vector{Float64}(grid)
        check(h, ccall((:seqj_set_grid, libseqj), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h.ptr, g64, length(g64)))
    end
    for m in metrics                                       # Distances.PeriodicEuclidean(periods): one period per row
        if m isa Distances.PeriodicEuclidean
            p64 = Vector{Float64}(m.periods)
            check(h, ccall((:seqj_set_periods, libseqj), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), h.ptr, p64, length(p64)))
        end
    end
    res = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:seqj_sequence, libseqj), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Int32}, Cint, Ptr{Int32}, Cint, Float64, Float64,
                Ptr{UInt8}, Ref{Ptr{Cvoid}}),
               h.ptr, A, M, N, mids, length(mids), scs, length(scs), ϵ, l2_threshold(metrics), C_NULL, res)
    check(h, st)
    r = res[]
    try
        EOSeg = Dict{Tuple,Any}()
        EOAlgScale = Dict{Tuple,Any}()
        g = 0
        for k in metrics, l in scales                                     # metric-major, as the reference loops
            met = Ref{Cint}(0); sc = Ref{Cint}(0); nch = Ref{Cint}(0); comp = Ref{Cint}(0)
            check(h, ccall((:seqj_result_group_info, libseqj), Cint,
                           (Ptr{Cvoid}, Cint, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}), r, g, met, sc, nch, comp))
            etas = Vector{Float64}(undef, nch[]); roots = Vector{Int32}(undef, nch[])
            pars = Matrix{Int32}(undef, N, nch[])
            check(h, ccall((:seqj_result_group_chunks, libseqj), Cint,
                           (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}), r, g, etas, roots, pars))
            EOSeg[(k, l)] = (Any[etas...], Any[bfs_digraph(pars[:, c], roots[c]) for c in 1:nch[]])
            eta = Ref{Float64}(0); root = Ref{Int32}(0); par = Vector{Int32}(undef, N)
            check(h, ccall((:seqj_result_group, libseqj), Cint,
                           (Ptr{Cvoid}, Cint, Ref{Float64}, Ref{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                           r, g, eta, root, par, C_NULL, C_NULL, C_NULL))
            EOAlgScale[(k, l)] = (eta[], bfs_digraph(par, root[]))
            tt = Ref{Float64}(0)                                          # device time of the group = the reference's `tt = @elapsed`
            check(h, ccall((:seqj_result_group_elapsed, libseqj), Cint, (Ptr{Cvoid}, Cint, Ref{Float64}), r, g, tt))
            @info "$(k) at scale $(l): η = $(@sprintf("%.4g", eta[])) ($(@sprintf("%.2g", tt[] / 1000))s)"   # src/sequencer.jl:260
            g += 1
        end
        eta = Ref{Float64}(0); root = Ref{Int32}(0)
        ord = Vector{Int32}(undef, N); par = Vector{Int32}(undef, N)
        mi = Vector{Int32}(undef, N - 1); mj = Vector{Int32}(undef, N - 1); mw = Vector{Float64}(undef, N - 1)
        check(h, ccall((:seqj_result_final, libseqj), Cint,
                       (Ptr{Cvoid}, Ref{Float64}, Ref{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
                       r, eta, root, ord, par, mi, mj, mw))
        nnz = Ref{Int64}(0)
        check(h, ccall((:seqj_result_D_nnz, libseqj), Cint, (Ptr{Cvoid}, Ref{Int64}), r, nnz))
        ti = Vector{Int32}(undef, nnz[]); tj = Vector{Int32}(undef, nnz[]); tv = Vector{Float64}(undef, nnz[])
        check(h, ccall((:seqj_result_D_triplets, libseqj), Cint,
                       (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}), r, ti, tj, tv))
        lo = Ref{Float64}(0); hi = Ref{Float64}(0)
        check(h, ccall((:seqj_result_input_range, libseqj), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), r, lo, hi))
        # D = inv.(Pw): +Inf off the edge union, ϵ on the diagonal after _mst's clamp (reference :269,360)
        Dm = fill(Inf, N, N)
        for i in 1:N; Dm[i, i] = ϵ; end
        for e in 1:nnz[]
            Dm[ti[e] + 1, tj[e] + 1] = tv[e]; Dm[tj[e] + 1, ti[e] + 1] = tv[e]
        end
        S = sparse(Int.(mi) .+ 1, Int.(mj) .+ 1, mw, N, N)
        mstD = SimpleWeightedGraph(S + S')
        ordr = Int.(ord) .+ 1                                             # 1-based column indices
        @info "Final: η = $(@sprintf("%.4g", eta[])) sequence = $(prettyp(ordr, 5))"
        return SequencerResult(EOSeg, EOAlgScale, Dm, mstD, eta[], ordr), lo[]
    finally
        ccall((:seqj_result_free, libseqj), Cvoid, (Ptr{Cvoid},), r)
    end
end

"Reference src/sequencer.jl:158-179; the column sample uses Julia's RNG (seed it for reproducibility)."
function autoscale(A::AbstractMatrix; scales = FIB, metrics = ALL_METRICS, grid = nothing, p = 0.1)
    M, N = size(A)
    Ŋ = round(Int, N * p, RoundDown)
    @assert Ŋ > 1 "Matrix A contains too few columns ($(N)) to sample with p = $(p). Use a larger p."
    sampidx = sort(rand(1:N, Ŋ))          # StatsBase.sample(..., ordered=true): with replacement, as the reference
    Asamp = convert(Matrix{Float64}, A[:, sampidx])
    maxscale = M ÷ 10
    @assert maxscale > 0 "Matrix A contains too few rows ($(M)) for autoscale. The minimum is 10 rows. It is recommended to set scale=(1,)."
    bestscale, max_η = 1, 0.0
    mets = metrics isa Union{PreMetric,Type} ? (metrics,) : tuple(metrics...)
    G = ensuregrid!(A, grid)
    for s in filter(i -> i < maxscale, FIB)
        r, _ = _sequence(Asamp, (s,), mets, G)
        if elong(r) > max_η
            max_η = elong(r); bestscale = s
        end
    end
    return bestscale
end

# ---- pair distances on explicit grids (reference src/distancemetrics.jl:82-87, 143-148) -----------
function _cdf_distance(mid::Integer, u, v, uw, vw)
    length(u) == length(uw) || error("u and u weights must have same length. Got u $(length(u)) and uw $(length(uw))")
    length(v) == length(vw) || error("v and v weights must have same length. Got v $(length(v)) and vw $(length(vw))")
    ndims(u) == 1 && ndims(v) == 1 || error("Only 1-d data vectors are supported.")
    h = handle()
    out = Ref{Float64}(0.0)
    check(h, ccall((:seqj_cdf_distance, libseqj), Cint,
                   (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}, Int64, Ref{Float64}),
                   h.ptr, Int32(mid), Vector{Float64}(u), Vector{Float64}(uw), length(u),
                   Vector{Float64}(v), Vector{Float64}(vw), length(v), out))
    out[]
end
"`emd(u, v, uw, vw)`: EMD between the weighted ECDFs on the grids `u`, `v` (reference :108)."
emd(u, v, uw, vw) = _cdf_distance(1, u, v, uw, vw)
"`emd(u, v)`: weights on the default unit grids `axes(u,1)`, `axes(v,1)` (reference :101)."
emd(u::AbstractVector, v = u) = _cdf_distance(1, collect(1.0:length(u)), collect(1.0:length(v)), u, v)
"`energy(u, v, uw, vw)` (reference :167) and `energy(u, v)` (:172)."
energy(u, v, uw, vw) = _cdf_distance(3, u, v, uw, vw)
energy(u::AbstractVector, v = u) = _cdf_distance(3, collect(1.0:length(u)), collect(1.0:length(v)), u, v)

# functor and four-argument forms of the types (reference src/distancemetrics.jl:46-50, 82-93, 143-162)
(m::EMD)(u::AbstractVector, v::AbstractVector = u) = m.grids === nothing ? emd(u, v) : emd(m.grids[1], m.grids[2], u, v)
(m::Energy)(u::AbstractVector, v::AbstractVector = u) = m.grids === nothing ? energy(u, v) : energy(m.grids[1], m.grids[2], u, v)
EMD(u::AbstractVector, v::AbstractVector, uw::AbstractVector, vw::AbstractVector) = emd(u, v, uw, vw)
Energy(u::AbstractVector, v::AbstractVector, uw::AbstractVector, vw::AbstractVector) = energy(u, v, uw, vw)

end # module
