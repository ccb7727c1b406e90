# KmerGMACuda.jl — the binding a KmerGMA.jl maintainer would add so that `ac_gma_testing!`, `Omn_KmerGMA!` and
# `exactMatch(query, path)` run on libkmergma_cuda (include/kmergma.h) while `findGenes(...)`,
# `findGenes_cluster_mode(...)` keep their keyword arguments and `Any[hit_vector, ...]` return value
# (src/API.jl:60-104, :161-226).  NOT EXECUTED in this repository's CI: Julia is not installed in the build
# image; the same call sequence is exercised through ctypes by kmergma.jl_b200/__init__.py and tests/.
#
# Drop-in use:  include("KmerGMACuda.jl") after `using KmerGMA`; the methods below replace the three operators.
module KmerGMACuda

using KmerGMA, FASTX, BioSequences

const LIB = get(ENV, "KMERGMA_CUDA_LIB", "libkmergma_cuda")

# ---- mirrors of the POD structs in include/kmergma.h -------------------------------------------------------
struct KgmaProfile                  # kgma_profile
    k::Int32; n_refs::Int32; window::Int64
    S::Ptr{Int32}; consensus::Ptr{UInt8}; consensus_len::Int32
    thr::Float64
end
struct KgmaScanParams               # kgma_scan_params
    mode::Int32; flags::UInt32; buff::Int64
    gap_open::Int32; gap_extend::Int32
    shard_index::Int32; shard_count::Int32; only_record::Int32; reserved::Int32
    strobe_s::Int32; strobe_w_min::Int32; strobe_w_max::Int32; strobe_q::Int32; score_threshold::Int64   # KGMA_MODE_STROBE only
end
struct KgmaHit                      # kgma_hit
    record::Int32; profile::Int32; cmi::Int64; first::Int64; last::Int64; genome_pos::Int64
    D::Int64; dist::Float64; align_score::Int64
    flags::UInt32; cigar_off::UInt32; cigar_len::UInt32; reserved::UInt32
end
struct KgmaAlignEvent               # kgma_align_event
    record::Int32; profile::Int32; cmi::Int64; align_score::Int64
    cigar_off::UInt32; cigar_len::UInt32; emitted::UInt32; reserved::UInt32
end
struct KgmaMatch; record::Int32; reserved::Int32; first::Int64; last::Int64; end

const F_ALIGN, F_DENSE, F_WANT_DISTS, F_WANT_CIGARS = UInt32(1), UInt32(2), UInt32(4), UInt32(8)

const F_RESIDENT = UInt32(32)

check(ctx, rc) = rc == 0 || error(unsafe_string(ccall((:kgma_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

# ---- one context per process and device: a kgma_ctx owns streams, page-locked staging blocks, device scratch and the device
# planes of the largest genome it has seen (a genome-sized cudaMalloc costs 0.1-0.4 s, a scan 1-20 ms), so it is created on
# first use and destroyed at exit -- never per call.  One context serves one Julia task at a time (calls are blocking).
const CONTEXTS = Dict{Int, Ptr{Cvoid}}()
const CTX_LOCK = ReentrantLock()

function context(device::Int = 0)
    lock(CTX_LOCK) do
        get!(CONTEXTS, device) do
            ctx = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:kgma_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ctx)
            rc == 0 || error(unsafe_string(ccall((:kgma_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))   # no B200 => error, never a CPU fallback
            ctx[]
        end
    end
end

# ---- genomes are ingested once per file (native parallel FASTA parse + 2-bit pack, kgma_genome_from_fasta) and kept, resident
# on the GPU, for as long as the file does not change: the second findGenes on the same genome pays ~1 ms, not the ingest.
const GENOMES = Dict{Tuple{String, Float64, Int64}, Ptr{Cvoid}}()

function genome(path::String, ctx)
    st = stat(path)
    key = (abspath(path), st.mtime, Int64(st.size))
    lock(CTX_LOCK) do
        g = get(GENOMES, key, C_NULL)
        if g == C_NULL
            for (k, old) in GENOMES          # one genome at a time: the context holds one set of device planes
                ccall((:kgma_genome_destroy, LIB), Cvoid, (Ptr{Cvoid},), old); delete!(GENOMES, k)
            end
            ref = Ref{Ptr{Cvoid}}(C_NULL)
            rc = ccall((:kgma_genome_from_fasta, LIB), Cint, (Cstring, Ref{Ptr{Cvoid}}), path, ref)
            rc == 0 || error(rc == -3 ? "KeyError: $path holds a symbol outside IUPAC DNA" : "cannot ingest $path ($rc)")
            g = ref[]
            check(ctx, ccall((:kgma_genome_make_resident, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, g))
            GENOMES[key] = g
        end
        g
    end
end

atexit() do
    for g in values(GENOMES); ccall((:kgma_genome_destroy, LIB), Cvoid, (Ptr{Cvoid},), g); end
    for c in values(CONTEXTS); ccall((:kgma_destroy, LIB), Cvoid, (Ptr{Cvoid},), c); end
end

identifier(g, r) = unsafe_string(ccall((:kgma_genome_identifier, LIB), Cstring, (Ptr{Cvoid}, Cint), g, r))

# view(seq, first:last) of record r, straight from the packed planes (N restored from the mask)
function subseq(g, r, rng::UnitRange)
    buf = Vector{UInt8}(undef, length(rng) + 1)
    ccall((:kgma_genome_get_seq, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{UInt8}), g, r, first(rng), last(rng), buf)
    return LongDNA{4}(String(buf[1:end-1]))
end

# A caller that already holds its records as LongDNA{4} (no FASTA file) hands BioSequences' 4-bit words over as they are
# (kgma_genome_append_bio4, no per-base Dict lookup).  A caller that packs to 2 bits itself can fill page-locked planes of the
# library in place instead: kgma_genome_create_pinned + kgma_genome_record_planes + kgma_genome_seal (zero copy).
function genome_from_records(records::Vector{FASTA.Record})
    g = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:kgma_genome_create, LIB), Cint, (Ref{Ptr{Cvoid}},), g)
    for record in records
        seq = getSeq(record)
        GC.@preserve seq begin
            rc = ccall((:kgma_genome_append_bio4, LIB), Cint, (Ptr{Cvoid}, Cstring, Cstring, Ptr{UInt64}, Int64),
                       g[], FASTA.identifier(record), FASTA.description(record), pointer(seq.data), length(seq))
            rc == 0 || error("kgma_genome_append_bio4 failed ($rc)")
        end
    end
    ccall((:kgma_genome_seal, LIB), Cint, (Ptr{Cvoid},), g[])
    return g[]
end

# refVec = S .* (1/N) (gen_ref_ws_cons) or S ./ N (cluster_ref_API): recover the integers the device needs
function ints_of(refVec::Vector{Float64})
    S = Vector{Int32}(undef, length(refVec)); n = Ref{Int32}(0)
    rc = ccall((:kgma_profile_from_kfv, LIB), Cint, (Ptr{Float64}, Int64, Int32, Ptr{Int32}, Ref{Int32}), refVec, length(refVec), 0, S, n)
    rc == 0 || error("refVec is not (k-mer count sums)/(family size)")
    return S, n[]
end

function scan!(resultVec, hit_loci_vec, dist_vecs, genome_path, refVecs, windowsizes, consensus_seqs, thrs, k, mode, buff,
               flags, gap_open, gap_extend; cluster::Bool, get_hit_loci::Bool, strobe = (0, 0, 0, 0), score_threshold::Int = 0,
               align_vec = nothing)
    ctx = context()
    g = genome(genome_path, ctx)
    Ss = [ints_of(Vector{Float64}(rv)) for rv in refVecs]
    cons = [Vector{UInt8}(string(c)) for c in consensus_seqs]
    GC.@preserve Ss cons begin
        profs = [KgmaProfile(k, Ss[i][2], windowsizes[i], pointer(Ss[i][1]), pointer(cons[i]), length(cons[i]), Float64(thrs[i])) for i in eachindex(Ss)]
        P = Ref(KgmaScanParams(mode, flags | F_RESIDENT, buff, gap_open, gap_extend, 0, 1, -1, 0, strobe..., score_threshold))
        res = Ref{Ptr{Cvoid}}(C_NULL)
        check(ctx, ccall((:kgma_scan, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{KgmaProfile}, Cint, Ref{KgmaScanParams}, Ref{Ptr{Cvoid}}),
                         ctx, g, profs, length(profs), P, res))
        try
            n = ccall((:kgma_result_n_hits, LIB), Int64, (Ptr{Cvoid},), res[])
            hits = unsafe_wrap(Array, ccall((:kgma_result_hits, LIB), Ptr{KgmaHit}, (Ptr{Cvoid},), res[]), n)
            for h in hits
                rng = h.first:h.last
                id = identifier(g, h.record)
                d = string(round(h.dist, digits = 2))                 # Julia's own rounding and printing
                header = cluster ?                                       # src/OmnGenomeMiner.jl:141-149 / src/Alignment.jl:71-78
                    id * " | Dist = " * d * " | KFV = $(h.profile) | MatchPos = $rng | GenomePos = $(h.genome_pos) | Len = " * string(length(rng)) :
                    id * " | dist = " * d * " | MatchPos = $rng | GenomePos = $(h.genome_pos) | Len = " * string(length(rng))
                push!(resultVec, FASTA.Record(header, subseq(g, h.record, rng)))
                get_hit_loci && push!(hit_loci_vec, h.first + h.genome_pos)
            end
            if align_vec !== nothing && (flags & F_WANT_CIGARS) != 0
                # do_return_align / get_aligns: (cigar, score) of every extension the reference pushes -- single mode one per hit
                # (Alignment.jl:46), cluster mode one per extension PERFORMED, rejected hits included (OmnGenomeMiner.jl:133)
                ops = ccall((:kgma_result_cigar_ops, LIB), Ptr{UInt8}, (Ptr{Cvoid},), res[])
                cnt = ccall((:kgma_result_cigar_counts, LIB), Ptr{Int32}, (Ptr{Cvoid},), res[])
                cig(off, len) = join(string(unsafe_load(cnt, off + t)) * Char(unsafe_load(ops, off + t)) for t in 1:len)
                if cluster
                    ne = ccall((:kgma_result_n_align_events, LIB), Int64, (Ptr{Cvoid},), res[])
                    evs = unsafe_wrap(Array, ccall((:kgma_result_align_events, LIB), Ptr{KgmaAlignEvent}, (Ptr{Cvoid},), res[]), ne)
                    for e in evs; push!(align_vec, (cigar = cig(e.cigar_off, e.cigar_len), score = e.align_score)); end
                else
                    for h in hits; push!(align_vec, (cigar = cig(h.cigar_off, h.cigar_len), score = h.align_score)); end
                end
            end
            if (flags & F_WANT_DISTS) != 0
                for q in eachindex(dist_vecs)
                    nd = ccall((:kgma_result_n_dists, LIB), Int64, (Ptr{Cvoid}, Cint), res[], q - 1)
                    append!(dist_vecs[q], unsafe_wrap(Array, ccall((:kgma_result_dists, LIB), Ptr{Float64}, (Ptr{Cvoid}, Cint), res[], q - 1), nd))
                end
            end
        finally
            ccall((:kgma_result_free, LIB), Cvoid, (Ptr{Cvoid},), res[])
        end
    end
end

# ---- the three operators, same keyword arguments as the reference (unused ones accepted and ignored) -------
function KmerGMA.ac_gma_testing!(; genome_path::String, refVec, consensus_refseq, k::Int64 = 6, windowsize::Int64 = 289,
        thr = 33.5, buff::Int64 = 50, mask = nothing, Nt_bits = nothing, ScaleFactor = nothing, do_align::Bool = true,
        result_align_vec = [], gap_open_score::Int = -69, gap_extend_score::Int = -1, do_return_dists::Bool = false,
        dist_vec = Float64[], do_return_align::Bool = false, get_hit_loci::Bool = false, hit_loci_vec = Int[],
        resultVec = FASTA.Record[])
    flags = (do_align ? F_ALIGN : UInt32(0)) | (do_return_dists ? F_WANT_DISTS : UInt32(0)) | ((do_align && do_return_align) ? F_WANT_CIGARS : UInt32(0))
    scan!(resultVec, hit_loci_vec, [dist_vec], genome_path, [collect(refVec)], [windowsize], [consensus_refseq], [thr], k, 0, buff,
          flags, gap_open_score, gap_extend_score; cluster = false, get_hit_loci = get_hit_loci, align_vec = result_align_vec)
end

function KmerGMA.Omn_KmerGMA!(; genome_path::String, refVecs, windowsizes, consensus_seqs, resultVec, k::Int = 6, ScaleFactor = nothing,
        mask = nothing, thr_vec = Float64[35, 31, 38, 34, 27, 27], buff::Int = 50, Nt_bits = nothing, align_hits::Bool = true,
        align_vec = [], gap_open_score::Int = -200, gap_extend_score::Int = -1, genome_pos::Int = 0, get_hit_loci::Bool = false,
        hit_loci_vec = Int[], get_aligns::Bool = false, do_return_dists::Bool = false, dist_vec_vec = [Float64[] for _ in windowsizes])
    flags = (align_hits ? F_ALIGN : UInt32(0)) | (do_return_dists ? F_WANT_DISTS : UInt32(0)) | ((align_hits && get_aligns) ? F_WANT_CIGARS : UInt32(0))
    scan!(resultVec, hit_loci_vec, dist_vec_vec, genome_path, refVecs, windowsizes, consensus_seqs, thr_vec, k, 1, buff,
          flags, gap_open_score, gap_extend_score; cluster = true, get_hit_loci = get_hit_loci, align_vec = align_vec)
end

# src/StrobemerGMA/StrobeGenomeMiner.jl:5-95 (KGMA_MODE_STROBE = 2; profile.k = w_max + s - 1, refVec over the 4^(2s) codes)
function KmerGMA.StrobeGMA!(; genome_path::String, refVec, consensus_refseq, s::Int = 2, w_min::Int = 3, w_max::Int = 5, q::Int = 5,
        windowsize::Int64 = 289, thr = 33.5, ScaleFactor = nothing, buff::Int64 = 50, do_align::Bool = true,
        score_model = AffineGapScoreModel(EDNAFULL, gap_open = -69, gap_extend = -5), score_threshold::Int = 0,
        do_return_dists::Bool = false, do_return_align::Bool = false, get_hit_loci::Bool = false, dist_vec = Float64[],
        result_align_vec = [], hit_loci_vec = Int[], genome_pos::Int = 0, resultVec = FASTA.Record[])
    flags = (do_align ? F_ALIGN : UInt32(0)) | (do_return_dists ? F_WANT_DISTS : UInt32(0))
    scan!(resultVec, hit_loci_vec, [dist_vec], genome_path, [collect(refVec)], [windowsize], [consensus_refseq], [thr], w_max + s - 1, 2, buff,
          flags, score_model.gap_open, score_model.gap_extend; cluster = false, get_hit_loci = get_hit_loci,
          strobe = (Int32(s), Int32(w_min), Int32(w_max), Int32(q)), score_threshold = score_threshold)
end

function KmerGMA.exactMatch(query, genome_path::String; overlap::Bool = true)
    ctx = context()
    g = genome(genome_path, ctx)
    q = Vector{UInt8}(string(query isa FASTA.Record ? getSeq(query) : query))
    out = Ref{Ptr{KgmaMatch}}(C_NULL); n = Ref{Int64}(0)
    check(ctx, ccall((:kgma_exact_match, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Int64, Cint, UInt32, Ref{Ptr{KgmaMatch}}, Ref{Int64}),
                     ctx, g, q, length(q), overlap, F_RESIDENT, out, n))
    identify = Dict{String, Vector{UnitRange{Int64}}}()
    for m in unsafe_wrap(Array, out[], n[])
        push!(get!(identify, identifier(g, m.record), UnitRange{Int64}[]), m.first:m.last)
    end
    n[] > 0 && ccall((:kgma_free, LIB), Cvoid, (Ptr{Cvoid},), out[])
    return isempty(identify) ? "no match" : identify      # src/ExactMatch.jl:116-120
end

end # module
