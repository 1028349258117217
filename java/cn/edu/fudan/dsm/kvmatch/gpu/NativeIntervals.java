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

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama binding of the library's host-only phase-1 tail (include/kvmatch_gpu.h: kvm_intervals_*, kvm_norm_intervals_*,
 * kvm_index_row_positions).  No GPU and no kvm_ctx behind these calls: they replace the boxed List&lt;Interval&gt; /
 * List&lt;NormInterval&gt; manipulation between index probing and verification.  Source only, never compiled in the build
 * image (no JDK there).  Interval lists live in MemorySegments: (left, right) int pairs plus one double per interval for
 * the RSM engines, 48-byte kvm_norm_interval structs for the cNSM engines.
 */
public final class NativeIntervals {

    private NativeIntervals() { }

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("kvmatch.gpu.lib", "libkvmatch_gpu.so"), Arena.global());

    /** struct kvm_norm_interval = K/common/NormInterval.java */
    public static final StructLayout NORM_INTERVAL = MemoryLayout.structLayout(
            JAVA_INT.withName("left"), JAVA_INT.withName("right"),
            JAVA_DOUBLE.withName("ex_lower"), JAVA_DOUBLE.withName("ex2_lower"),
            JAVA_DOUBLE.withName("ex_upper"), JAVA_DOUBLE.withName("ex2_upper"),
            JAVA_LONG.withName("beta_partitions"));

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }

    // int kvm_intervals_sort_merge(lr, eps, k, mode, lr_out, eps_out, cap, k_out, cnt_disjoint, cnt_offsets)
    private static final MethodHandle SORT_MERGE = fn("kvm_intervals_sort_merge", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    // int kvm_intervals_intersect(cs_lr, cs_eps, k1, csi_lr, csi_eps, k2, eps2, delta_w, lr_out, eps_out, cap, k_out, min_eps)
    private static final MethodHandle INTERSECT = fn("kvm_intervals_intersect", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, JAVA_LONG, JAVA_DOUBLE, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    // int kvm_intervals_first_segment(lr, eps, k, order, w0, length, n, delta_w, lr_out, eps_out, cap, k_out, min_eps)
    private static final MethodHandle FIRST_SEGMENT = fn("kvm_intervals_first_segment", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    // int kvm_norm_intervals_sort_merge(in, k, mode, out, cap, k_out, cnt_disjoint, cnt_offsets)
    private static final MethodHandle NORM_SORT_MERGE = fn("kvm_norm_intervals_sort_merge", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    // int kvm_norm_intervals_intersect(cs, k1, csi, k2, pre_length, w0, query_length, mean_q, std_q, alpha, beta, delta_w, dtw, out, cap, k_out)
    private static final MethodHandle NORM_INTERSECT = fn("kvm_norm_intervals_intersect", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE,
            JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    // int kvm_norm_intervals_first_segment(in, k, order, w0, length, n, delta_w, out, cap, k_out)
    private static final MethodHandle NORM_FIRST_SEGMENT = fn("kvm_norm_intervals_first_segment", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    // int kvm_index_row_positions(row, row_bytes, lr_out, cap, k_out)
    private static final MethodHandle ROW_POSITIONS = fn("kvm_index_row_positions", FunctionDescriptor.of(JAVA_INT,
            ADDRESS, JAVA_LONG, ADDRESS, JAVA_LONG, ADDRESS));

    /** A cNSM interval list: `size` kvm_norm_interval structs at the start of `seg`. */
    public record NormList(MemorySegment seg, long size) { }

    private static void check(int rc, String what) throws IOException {
        if (rc != 0) throw new IOException(what + " failed: code " + rc);
    }

    /** mode 0 / 1 / 2 = sortButNotMergeIntervals / ...AndCount / sortAndMergeIntervals (K/NormQueryEngine.java:788-896).
     *  counts (may be null) receives {cntDisjointIntervals, cntOffsets}. */
    public static NormList normSortMerge(Arena arena, NormList in, int mode, long[] counts) throws IOException {
        MemorySegment out = arena.allocate(NORM_INTERVAL, Math.max(in.size(), 1));
        MemorySegment k = arena.allocate(JAVA_LONG), cd = arena.allocate(JAVA_LONG), co = arena.allocate(JAVA_LONG);
        try {
            check((int) NORM_SORT_MERGE.invokeExact(in.seg(), in.size(), mode, out, Math.max(in.size(), 1), k, cd, co), "kvm_norm_intervals_sort_merge");
        } catch (IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
        if (counts != null) {
            counts[0] = cd.get(JAVA_LONG, 0);
            counts[1] = co.get(JAVA_LONG, 0);
        }
        return new NormList(out, k.get(JAVA_LONG, 0));
    }

    /** CS ∩ CS_i with the beta-partition and variance filters (K/NormQueryEngine.java:333-397; dtw: NormQueryEngineDtw.java:349-425). */
    public static NormList normIntersect(Arena arena, NormList cs, NormList csi, int preLength, int w0, int queryLength, double meanQ,
                                         double stdQ, double alpha, double beta, int deltaW, boolean dtw) throws IOException {
        long cap = Math.max(cs.size() + csi.size(), 1);
        MemorySegment out = arena.allocate(NORM_INTERVAL, cap);
        MemorySegment k = arena.allocate(JAVA_LONG);
        try {
            check((int) NORM_INTERSECT.invokeExact(cs.seg(), cs.size(), csi.seg(), csi.size(), preLength, w0, queryLength, meanQ, stdQ,
                    alpha, beta, deltaW, dtw ? 1 : 0, out, cap, k), "kvm_norm_intervals_intersect");
        } catch (IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
        return new NormList(out, k.get(JAVA_LONG, 0));
    }

    /** The first segment's positions clamped to window starts inside the series (K/NormQueryEngine.java:313-332). */
    public static NormList normFirstSegment(Arena arena, NormList in, int order, int w0, int length, int n, int deltaW) throws IOException {
        MemorySegment out = arena.allocate(NORM_INTERVAL, Math.max(in.size(), 1));
        MemorySegment k = arena.allocate(JAVA_LONG);
        try {
            check((int) NORM_FIRST_SEGMENT.invokeExact(in.seg(), in.size(), order, w0, length, n, deltaW, out, Math.max(in.size(), 1), k),
                    "kvm_norm_intervals_first_segment");
        } catch (IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
        return new NormList(out, k.get(JAVA_LONG, 0));
    }

    /** (left, right) pairs of one index row from its compact bytes, the 8-byte key excluded (IndexNode.parseBytesCompact). */
    public static int[] rowPositions(Arena arena, MemorySegment row, long rowBytes) throws IOException {
        long cap = Math.max(rowBytes / 2, 1);
        MemorySegment out = arena.allocate(JAVA_INT, 2 * cap);
        MemorySegment k = arena.allocate(JAVA_LONG);
        try {
            check((int) ROW_POSITIONS.invokeExact(row, rowBytes, out, cap, k), "kvm_index_row_positions");
        } catch (IOException e) {
            throw e;
        } catch (Throwable t) {
            throw new IOException(t);
        }
        return out.asSlice(0, 8 * k.get(JAVA_LONG, 0)).toArray(JAVA_INT);
    }

    // The RSM forms (SORT_MERGE, INTERSECT, FIRST_SEGMENT) take two parallel segments — int pairs and one double per
    // interval — and are called the same way; see INTEGRATION.md, "Phase-1 tail".
    static MethodHandle rsmSortMerge() { return SORT_MERGE; }
    static MethodHandle rsmIntersect() { return INTERSECT; }
    static MethodHandle rsmFirstSegment() { return FIRST_SEGMENT; }
}
