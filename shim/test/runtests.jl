# First thing to run on a box that has Julia + Springsteel + a B200: Springsteel's own transforms against
# libscythe_b200 on the same seeded field (the parity pin this repository cannot execute: no Julia in its image).
using Test, Random
using Springsteel
include(joinpath(@__DIR__, "..", "src", "ScytheB200.jl"))
using .ScytheB200

@testset "transforms vs Springsteel" begin
    for (geom, kw) in (("R", (;)), ("RL", (;)), ("RZ", (zmin = 0.0, zmax = 5.0, zDim = 12)), ("RLZ", (zmin = 0.0, zmax = 5.0, zDim = 10)))
        gp = GridParameters(; geometry = geom, xmin = 0.0, xmax = 10.0, num_cells = 6,
                            BCL = Dict("u" => CubicBSpline.R1T1), BCR = Dict("u" => CubicBSpline.R0),
                            BCB = Dict("u" => Chebyshev.R0), BCT = Dict("u" => Chebyshev.R0), vars = Dict("u" => 1), kw...)
        ref = Springsteel.createGrid(gp)
        gpu = ScytheB200.createGrid(gp)
        Random.seed!(1)
        u = randn(size(ref.physical, 1))
        ref.physical[:, 1, 1] .= u
        gpu.physical[:, 1, 1] .= u
        @test maximum(abs.(vec(ScytheB200.getGridpoints(gpu)) .- vec(Springsteel.getGridpoints(ref)))) < 1e-12 * 10
        Springsteel.spectralTransform!(ref)
        ScytheB200.spectralTransform!(gpu)
        @test maximum(abs.(gpu.spectral .- ref.spectral)) <= 1e-12 * maximum(abs.(ref.spectral))
        Springsteel.gridTransform!(ref)
        ScytheB200.gridTransform!(gpu)
        for d in 1:size(ref.physical, 3)
            @test maximum(abs.(gpu.physical[:, :, d] .- ref.physical[:, :, d])) <= 1e-12 * max(maximum(abs.(ref.physical[:, :, d])), 1e-300)
        end
    end
end
