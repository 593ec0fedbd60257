/*
 * DumpGolden.java -- golden-vector harness for the REAL reference (c0der1337/Neural_Tensor_Network_For_KB_Completion).
 *
 * NOT COMPILED IN THIS REPOSITORY: the build image has no JDK and the reference's jars (ND4j 0.0.3.5.5.3-SNAPSHOT,
 * jblas 1.2.3, umass-nlp, jmatio; /root/reference/.classpath:6-20) are not vendored.  A maintainer of the reference
 * drops this file next to src/Run_NTN.java and runs, on a machine with a JVM and the jars:
 *
 *     javac -cp "lib/*" -d bin src/*.java tools/DumpGolden.java
 *     java  --add-opens java.base/java.lang=ALL-UNNAMED -Xmx4G -cp "bin:lib/*" DumpGolden <data_path> <out_dir> [seed]
 *
 * It performs ONE cost+gradient evaluation exactly as Run_NTN.main does (Run_NTN.java:55-56,72-83,90-107,115) --
 * same constructor arguments, same loaders, NTN.computeAt(theta) on a fresh batch job -- and writes everything the
 * evaluation read and produced to <out_dir> in the layout tests/test_reference_dump.py loads:
 *
 *     manifest.json                d, k, R, Ne, Nw, batchSize, corruptSize, lambda, activation, P, n_columns, seed
 *     entities.txt, words.txt      entity names / vocabulary in index order (one per line)
 *     wv_loaded.f32                DataFactory.getWordVectorMaxtrixLoaded(), d x Nw row-major, little-endian float32
 *     bwd_words.i32                per entity: entityLength(i), n, getWordIndexes(i)[0..n)      (DataFactory.java:583-595,646-684)
 *     batch.i32                    n_columns x {e1, rel, e2, e3}, relation by relation in batch order (DataFactory.java:89-98)
 *     sides.u8                     R bytes: 1 = this evaluation corrupted the tail of relation r (NTN.java:146-150)
 *     theta.f64, grad.f64          P doubles each: the argument and the returned gradient (NTN.java:93,442)
 *     cost.f64                     the returned cost
 *
 * The two unseeded random sources of the evaluation are pinned without touching the reference's sources:
 *   - the corrupt entities (new Random() at DataFactory.java:119) are read back from the generated batch job;
 *   - Math.random() (NTN.java:146) is replaced, through reflection on java.lang.Math's holder class, by a Random with a
 *     known seed, and the same stream is replayed here: one nextDouble() per relation that has batch columns
 *     (NTN.java:120 skips empty relations before the draw), in relation order.
 * tests/test_reference_dump.py feeds these inputs to oracle/ntn_ref.py (restatement check: this is what would PIN the
 * oracle) and to the CUDA path (parity check) and compares cost and gradient at 1e-6 + 1e-4|ref|.
 */
import java.io.DataOutputStream;
import java.io.File;
import java.io.FileOutputStream;
import java.io.PrintWriter;
import java.lang.reflect.Field;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.Random;

import org.nd4j.linalg.api.buffer.DataBuffer;
import org.nd4j.linalg.api.ndarray.INDArray;
import org.nd4j.linalg.factory.Nd4j;

import edu.umass.nlp.utils.IPair;

public class DumpGolden {
    static void writeLE(String path, ByteBuffer b) throws Exception {
        try (FileOutputStream f = new FileOutputStream(path)) { f.write(b.array(), 0, b.position()); }
    }
    static ByteBuffer le(long bytes) { return ByteBuffer.allocate((int) bytes).order(ByteOrder.LITTLE_ENDIAN); }

    /** Replace the generator behind Math.random() (java.lang.Math$RandomNumberGeneratorHolder.randomNumberGenerator). */
    static void seedMathRandom(long seed) throws Exception {
        Class<?> holder = Class.forName("java.lang.Math$RandomNumberGeneratorHolder");
        Field f = holder.getDeclaredField("randomNumberGenerator");
        f.setAccessible(true);
        try {   // final static: clear the modifier where the JDK still allows it (<= 11); newer JDKs need --add-opens + Unsafe
            Field m = Field.class.getDeclaredField("modifiers");
            m.setAccessible(true);
            m.setInt(f, f.getModifiers() & ~java.lang.reflect.Modifier.FINAL);
            f.set(null, new Random(seed));
        } catch (NoSuchFieldException e) {
            Field uf = Class.forName("sun.misc.Unsafe").getDeclaredField("theUnsafe");
            uf.setAccessible(true);
            sun.misc.Unsafe u = (sun.misc.Unsafe) uf.get(null);
            u.putObject(u.staticFieldBase(f), u.staticFieldOffset(f), new Random(seed));
        }
    }

    public static void main(String[] args) throws Exception {
        String dataPath = args[0], out = args[1];
        long seed = args.length > 2 ? Long.parseLong(args[2]) : 7L;
        new File(out).mkdirs();
        Nd4j.dtype = DataBuffer.FLOAT;                                   // Run_NTN.java:55
        org.nd4j.linalg.api.ndarray.BaseNDArray.class.getName();         // (ENFORCE_NUMERICAL_STABILITY is set below)
        Nd4j.ENFORCE_NUMERICAL_STABILITY = true;                         // Run_NTN.java:56

        // Run_NTN.java:72-83
        int batchSize = 20000, numWVdimensions = 100, sliceSize = 3, corrupt_size = 10;
        String activation_function = "tanh";
        float lamda = 0.0001F;

        // Run_NTN.java:90-100 (argument order as in the reference, both flags false)
        DataFactory tbj = DataFactory.getInstance(batchSize, corrupt_size, numWVdimensions, false, false);
        tbj.loadEntitiesFromSocherFile(dataPath + "entities.txt");
        tbj.loadRelationsFromSocherFile(dataPath + "relations.txt");
        tbj.loadTrainingDataTripplesE1rE2(dataPath + "train.txt");
        tbj.loadDevDataTripplesE1rE2Label(dataPath + "dev.txt");
        tbj.loadTestDataTripplesE1rE2Label(dataPath + "test.txt");
        tbj.loadWordVectorsFromMatFile(dataPath + "initEmbed.mat", false);

        // Run_NTN.java:103-107
        NTN t = new NTN(numWVdimensions, tbj.getNumOfentities(), tbj.getNumOfRelations(), tbj.getNumOfWords(), batchSize,
                        sliceSize, activation_function, tbj, lamda);
        t.connectDatafactory(tbj);
        double[] theta = t.getTheta_inital().data().asDouble();

        tbj.generateNewTrainingBatchJob();                               // Run_NTN.java:115
        int R = tbj.getNumOfRelations(), Ne = tbj.getNumOfentities(), Nw = tbj.getNumOfWords();

        // the batch job, relation by relation in batch order, and the corrupt sides this evaluation will draw
        ArrayList<int[]> cols = new ArrayList<int[]>();
        byte[] sides = new byte[R];
        Random shadow = new Random(seed);
        for (int r = 0; r < R; r++) {
            ArrayList<Tripple> tr = tbj.getBatchJobTripplesOfRelation(r);
            for (Tripple x : tr)
                cols.add(new int[] {x.getIndex_entity1(), x.getIndex_relation(), x.getIndex_entity2(), x.getIndex_entity3_corrupt()});
            if (tr.size() > 0) sides[r] = (byte) (shadow.nextDouble() > 0.5 ? 1 : 0);   // NTN.java:120,146
        }
        seedMathRandom(seed);
        IPair<Double, double[]> res = t.computeAt(theta);                // NTN.java:93
        double cost = res.getFirst();
        double[] grad = res.getSecond();

        try (PrintWriter w = new PrintWriter(out + "/manifest.json")) {
            w.printf("{\"d\": %d, \"k\": %d, \"R\": %d, \"Ne\": %d, \"Nw\": %d, \"batchSize\": %d, \"corruptSize\": %d, "
                     + "\"lambda\": %s, \"activation\": \"%s\", \"P\": %d, \"n_columns\": %d, \"seed\": %d, "
                     + "\"wv_source\": \"frozen\", \"producer\": \"tools/DumpGolden.java on the reference's own classes\"}%n",
                     numWVdimensions, sliceSize, R, Ne, Nw, batchSize, corrupt_size, Float.toString(lamda), activation_function,
                     theta.length, cols.size(), seed);
        }
        // names: the reference keeps them in private maps; the loaders read them from the same files in index order,
        // so the test re-reads entities.txt / the .mat vocabulary itself.  What IS dumped is what the reference derived:
        ByteBuffer bw = le(8L * Ne + 4L * 8 * Ne);
        for (int i = 0; i < Ne; i++) {
            int[] wi = tbj.getWordIndexes(i);
            bw.putInt(tbj.entityLength(i)).putInt(wi.length);
            for (int x : wi) bw.putInt(x);
        }
        writeLE(out + "/bwd_words.i32", bw);
        INDArray wv = tbj.getWordVectorMaxtrixLoaded();                  // d x Nw
        ByteBuffer bv = le(4L * wv.rows() * wv.columns());
        for (int i = 0; i < wv.rows(); i++) for (int j = 0; j < wv.columns(); j++) bv.putFloat(wv.getFloat(i, j));
        writeLE(out + "/wv_loaded.f32", bv);
        ByteBuffer bb = le(16L * cols.size());
        for (int[] c : cols) for (int x : c) bb.putInt(x);
        writeLE(out + "/batch.i32", bb);
        try (FileOutputStream f = new FileOutputStream(out + "/sides.u8")) { f.write(sides); }
        ByteBuffer bt = le(8L * theta.length), bg = le(8L * grad.length), bc = le(8);
        for (double x : theta) bt.putDouble(x);
        for (double x : grad) bg.putDouble(x);
        bc.putDouble(cost);
        writeLE(out + "/theta.f64", bt);
        writeLE(out + "/grad.f64", bg);
        writeLE(out + "/cost.f64", bc);
        System.out.println("DumpGolden: cost " + cost + ", P " + theta.length + ", columns " + cols.size() + " -> " + out
                           + "  (copy " + dataPath + "entities.txt and the .mat vocabulary order next to it as entities.txt / words.txt)");
    }
}
