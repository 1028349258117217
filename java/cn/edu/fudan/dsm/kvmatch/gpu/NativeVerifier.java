package cn.edu.fudan.dsm.kvmatch.gpu;

import java.io.IOException;
import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;
import java.util.List;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama (java.lang.foreign, JDK 22+) binding of libkvmatch_gpu.so (include/kvmatch_gpu.h).
 * Source only: no JDK exists in the build image, so this file has never been compiled there.
 * One instance = one kvm_ctx = one GPU holding the series; not re-entrant, like the engines themselves.
 * A non-zero return code becomes an IOException (the engines' query() already declares it); there is no CPU fallback.
 */
public final class NativeVerifier implements AutoCloseable {

    /** One answer: 1-based offset and sqrt(dist^2), as Pair<Integer, Double> in the engines. */
    public record Answer(int offset, double distance) { }

    /** Step-1 output of IndexBuilder for one window width. */
    public record Runs(double[] keys, int[] first, int[] last) { }

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("kvmatch.gpu.lib", "libkvmatch_gpu.so"), Arena.global());

    // struct kvm_result (include/kvmatch_gpu.h)
    private static final StructLayout RESULT = MemoryLayout.structLayout(
            JAVA_LONG.withName("count"), ADDRESS.withName("offsets"), ADDRESS.withName("distances"),
            JAVA_LONG.withName("cnt_candidate"), JAVA_LONG.withName("n_verified"), JAVA_LONG.withName("s_total"),
            JAVA_LONG.withName("n_gate_pass"), JAVA_LONG.withName("n_lb_pass"), JAVA_LONG.withName("n_exact"),
            JAVA_DOUBLE.withName("kernel_ms"), MemoryLayout.sequenceLayout(4, JAVA_DOUBLE).withName("stage_ms"),
            JAVA_INT.withName("n_launches"), JAVA_INT.withName("h2d_bytes"),
            JAVA_LONG.withName("n_rewalked"), JAVA_LONG.withName("n_chains_rewalked"), JAVA_LONG.withName("n_dtw_cells"));
    // struct kvm_runs
    private static final StructLayout RUNS = MemoryLayout.structLayout(
            JAVA_LONG.withName("count"), ADDRESS.withName("keys"), ADDRESS.withName("first"), ADDRESS.withName("last"),
            JAVA_DOUBLE.withName("kernel_ms"), JAVA_INT.withName("n_launches"), JAVA_INT.withName("reserved"));

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }

    private static final MethodHandle CREATE = fn("kvm_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle DESTROY = fn("kvm_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LAST_ERROR = fn("kvm_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle LOAD_FILE = fn("kvm_load_series_file",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_LONG));
    private static final MethodHandle VERIFY_ED = fn("kvm_verify_ed",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_DOUBLE, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle VERIFY_CNSM_ED = fn("kvm_verify_cnsm_ed",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS,
                    JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle VERIFY_DTW = fn("kvm_verify_dtw",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT,
                    ADDRESS));
    private static final MethodHandle VERIFY_CNSM_DTW = fn("kvm_verify_cnsm_dtw",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE,
                    ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle SCAN_UCR_DTW = fn("kvm_scan_ucr_dtw",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE,
                    ADDRESS));
    private static final MethodHandle SCAN_UCR_ED = fn("kvm_scan_ucr_ed",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS));
    // struct kvm_index_info
    private static final StructLayout INDEX_INFO = MemoryLayout.structLayout(
            JAVA_LONG.withName("file_bytes"), JAVA_LONG.withName("n_runs"), JAVA_LONG.withName("n_intervals"),
            JAVA_LONG.withName("n_offsets"), JAVA_INT.withName("n_rows_step1"), JAVA_INT.withName("n_rows"),
            JAVA_DOUBLE.withName("kernel_ms"), JAVA_DOUBLE.withName("host_ms"));
    private static final MethodHandle BUILD_INDEX_FILE = fn("kvm_build_index_file",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle WINDOW_MEAN_RUNS = fn("kvm_window_mean_runs",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));

    private final MemorySegment ctx;
    public long lastCntCandidate;

    /** Loads files/data-N (big-endian doubles, K/DataGenerator.java:102-113) onto GPU `device`. */
    public NativeVerifier(int device, String dataFile, long n) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            int rc = (int) CREATE.invokeExact(out, device);
            if (rc != 0) throw new IOException("kvm_create: " + error(MemorySegment.NULL) + " (" + rc + ")");
            ctx = out.get(ADDRESS, 0);
            rc = (int) LOAD_FILE.invokeExact(ctx, a.allocateFrom(dataFile), n, 1L, n);
            check(rc);
        } catch (IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
    }

    private String error(MemorySegment c) throws Throwable {
        MemorySegment s = (MemorySegment) LAST_ERROR.invokeExact(c);
        return s.reinterpret(512).getString(0);
    }

    private void check(int rc) throws Throwable {
        if (rc != 0) throw new IOException("libkvmatch_gpu: " + error(ctx) + " (" + rc + ")");
    }

    private static MemorySegment doubles(Arena a, List<Double> q) {
        MemorySegment s = a.allocate(JAVA_DOUBLE, q.size());
        for (int i = 0; i < q.size(); i++) s.setAtIndex(JAVA_DOUBLE, i, q.get(i));
        return s;
    }

    /** lr = {left0, right0, left1, right1, ...} of the merged validPositions. */
    private static MemorySegment ints(Arena a, int[] lr) {
        return a.allocateFrom(JAVA_INT, lr);
    }

    private List<Answer> take(MemorySegment res) {
        long count = res.get(JAVA_LONG, 0);
        lastCntCandidate = res.get(JAVA_LONG, 24);
        MemorySegment off = res.get(ADDRESS, 8).reinterpret(4 * count);
        MemorySegment dist = res.get(ADDRESS, 16).reinterpret(8 * count);
        List<Answer> answers = new ArrayList<>((int) count);
        for (long i = 0; i < count; i++) answers.add(new Answer(off.getAtIndex(JAVA_INT, i), dist.getAtIndex(JAVA_DOUBLE, i)));
        return answers;  // ascending offset = the reference's scan order; callers stable-sort by distance
    }

    public List<Answer> verifyEd(List<Double> q, double epsilon, int[] lr, int shift) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(RESULT);
            check((int) VERIFY_ED.invokeExact(ctx, doubles(a, q), q.size(), epsilon, ints(a, lr), lr.length / 2, shift, res));
            return take(res);
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    public List<Answer> verifyCnsmEd(List<Double> q, double epsilon, double alpha, double beta, int[] lr, int shift)
            throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(RESULT);
            check((int) VERIFY_CNSM_ED.invokeExact(ctx, doubles(a, q), q.size(), epsilon, alpha, beta, ints(a, lr),
                    lr.length / 2, shift, res));
            return take(res);
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    public List<Answer> verifyDtw(List<Double> q, double epsilon, int rho, int[] lr, int shift) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(RESULT);
            check((int) VERIFY_DTW.invokeExact(ctx, doubles(a, q), q.size(), epsilon, rho, ints(a, lr), lr.length / 2, shift, res));
            return take(res);
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    /** SingleIndexBuilder.run() for one window width: writes the index file, returns the number of rows. */
    public int buildIndexFile(int w, String path) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment info = a.allocate(INDEX_INFO);
            check((int) BUILD_INDEX_FILE.invokeExact(ctx, w, a.allocateFrom(path), info));
            return info.get(JAVA_INT, INDEX_INFO.byteOffset(MemoryLayout.PathElement.groupElement("n_rows")));
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    /** Index-free scan, UcrDtwQueryExecutor semantics (0-based offsets). */
    public List<Answer> scanUcrDtw(List<Double> q, double epsilon, int rho, double alpha, double beta) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(RESULT);
            check((int) SCAN_UCR_DTW.invokeExact(ctx, doubles(a, q), q.size(), epsilon, rho, alpha, beta, res));
            return take(res);
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    /** Index-free scan, UcrEdQueryExecutor semantics (one never-reset statistics chain, 1-based offsets). */
    public List<Answer> scanUcrEd(List<Double> q, double epsilon, double alpha, double beta) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(RESULT);
            check((int) SCAN_UCR_ED.invokeExact(ctx, doubles(a, q), q.size(), epsilon, alpha, beta, res));
            return take(res);
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    public List<Answer> verifyCnsmDtw(List<Double> q, double epsilon, int rho, double alpha, double beta, int[] lr, int shift)
            throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(RESULT);
            check((int) VERIFY_CNSM_DTW.invokeExact(ctx, doubles(a, q), q.size(), epsilon, rho, alpha, beta, ints(a, lr),
                    lr.length / 2, shift, res));
            return take(res);
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    public Runs windowMeanRuns(int w) throws IOException {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment res = a.allocate(RUNS);
            check((int) WINDOW_MEAN_RUNS.invokeExact(ctx, w, res));
            int count = (int) res.get(JAVA_LONG, 0);
            return new Runs(res.get(ADDRESS, 8).reinterpret(8L * count).toArray(JAVA_DOUBLE),
                    res.get(ADDRESS, 16).reinterpret(4L * count).toArray(JAVA_INT),
                    res.get(ADDRESS, 24).reinterpret(4L * count).toArray(JAVA_INT));
        } catch (IOException e) { throw e; } catch (Throwable t) { throw new IOException(t); }
    }

    @Override
    public void close() {
        try { DESTROY.invokeExact(ctx); } catch (Throwable ignored) { }
    }
}
