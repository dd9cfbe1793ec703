// MecanoHarness -- runs Mecano's OWN calculators (the JVM reference) on a system and states written by this repository, so that
//   (1) the GPU results / the C oracle can be compared with the Java implementation value by value ("parity unpinned" -> pinned), and
//   (2) the CPU baseline beside the GPU number is the real multithreaded JVM path (one cloned system + calculator per thread,
//       MultiBodySystemFactories.cloneMultiBodySystem, as SURVEY.md 8(d) specifies), not the C restatement.
//
// STATUS: source only.  No JDK exists in the build image or on the GPU boxes, so this file has never been compiled or run; the
// Mecano / Euclid / EJML calls were written against the reference sources (file:line cited at each use).  See README.md here.
//
//   javac -cp "mecano.jar:euclid.jar:euclid-frame.jar:euclid-geometry.jar:ejml-core.jar:ejml-ddense.jar" MecanoHarness.java
//   java  -cp ".:<same jars>" MecanoHarness dump  exchange.bin results.bin          # results of every state, for the cross-check
//   java  -cp ".:<same jars>" MecanoHarness bench exchange.bin [threads] [seconds]  # states/s of RNEA + ABA + CRBA, all cores
//
// exchange.bin is written by scripts/java_exchange.py (little endian):
//   int32 magic 0x4D423258, nb, nv, nq, n;  double gravity[3]
//   nb x { int32 jointType (0 revolute, 1 prismatic, 2 six-DoF), int32 parent (-1 = root body), double axis[3], offsetR[9],
//          offsetP[3], comR[9], comP[3], inertia[9], mass }            joints in JointMatrixIndexProvider (depth-first) order
//   double q[nq][n], qd[nv][n], qdd[nv][n], tau[nv][n]                    row-major, state-minor (the layout of the C ABI)
// results.bin: double tauID[nv][n], qddFD[nv][n], M[nv*nv][n]            same layout
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.channels.FileChannel;
import java.nio.file.Paths;
import java.nio.file.StandardOpenOption;
import java.util.ArrayList;
import java.util.List;
import java.util.concurrent.atomic.AtomicLong;

import org.ejml.data.DMatrixRMaj;

import us.ihmc.euclid.matrix.Matrix3D;
import us.ihmc.euclid.referenceFrame.ReferenceFrame;
import us.ihmc.euclid.transform.RigidBodyTransform;
import us.ihmc.euclid.tuple3D.Vector3D;
import us.ihmc.mecano.algorithms.CompositeRigidBodyMassMatrixCalculator;
import us.ihmc.mecano.algorithms.ForwardDynamicsCalculator;
import us.ihmc.mecano.algorithms.InverseDynamicsCalculator;
import us.ihmc.mecano.multiBodySystem.PrismaticJoint;
import us.ihmc.mecano.multiBodySystem.RevoluteJoint;
import us.ihmc.mecano.multiBodySystem.RigidBody;
import us.ihmc.mecano.multiBodySystem.SixDoFJoint;
import us.ihmc.mecano.multiBodySystem.interfaces.JointBasics;
import us.ihmc.mecano.multiBodySystem.interfaces.MultiBodySystemBasics;
import us.ihmc.mecano.multiBodySystem.interfaces.RigidBodyBasics;
import us.ihmc.mecano.tools.JointStateType;
import us.ihmc.mecano.tools.MultiBodySystemTools;

public class MecanoHarness
{
   static final int MAGIC = 0x4D423258;

   static class Exchange
   {
      int nb, nv, nq, n;
      double[] gravity = new double[3];
      int[] jointType, parent;
      double[][] axis, offsetR, offsetP, comR, comP, inertia;
      double[] mass;
      double[] q, qd, qdd, tau; // [rows][n] row-major
   }

   static double[] doubles(ByteBuffer b, int count)
   {
      double[] d = new double[count];
      for (int i = 0; i < count; i++)
         d[i] = b.getDouble();
      return d;
   }

   static Exchange read(String path) throws IOException
   {
      try (FileChannel ch = FileChannel.open(Paths.get(path), StandardOpenOption.READ))
      {
         ByteBuffer b = ch.map(FileChannel.MapMode.READ_ONLY, 0, ch.size()).order(ByteOrder.LITTLE_ENDIAN);
         if (b.getInt() != MAGIC)
            throw new IOException("not an exchange file: " + path);
         Exchange x = new Exchange();
         x.nb = b.getInt(); x.nv = b.getInt(); x.nq = b.getInt(); x.n = b.getInt();
         x.gravity = doubles(b, 3);
         x.jointType = new int[x.nb]; x.parent = new int[x.nb]; x.mass = new double[x.nb];
         x.axis = new double[x.nb][]; x.offsetR = new double[x.nb][]; x.offsetP = new double[x.nb][];
         x.comR = new double[x.nb][]; x.comP = new double[x.nb][]; x.inertia = new double[x.nb][];
         for (int i = 0; i < x.nb; i++)
         {
            x.jointType[i] = b.getInt(); x.parent[i] = b.getInt();
            x.axis[i] = doubles(b, 3); x.offsetR[i] = doubles(b, 9); x.offsetP[i] = doubles(b, 3);
            x.comR[i] = doubles(b, 9); x.comP[i] = doubles(b, 3); x.inertia[i] = doubles(b, 9);
            x.mass[i] = b.getDouble();
         }
         x.q = doubles(b, x.nq * x.n); x.qd = doubles(b, x.nv * x.n); x.qdd = doubles(b, x.nv * x.n); x.tau = doubles(b, x.nv * x.n);
         return x;
      }
   }

   static RigidBodyTransform transform(double[] R, double[] p)
   {
      RigidBodyTransform t = new RigidBodyTransform();
      // RotationMatrixBasics.set(Matrix3DReadOnly) checks / re-normalises the matrix; setUnsafe(m00, ..., m22) is the unchecked form
      t.getRotation().set(new Matrix3D(R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8]));
      t.getTranslation().set(p[0], p[1], p[2]);
      return t;
   }

   /** Builds the Mecano system: joints are created in the depth-first order of the file, so JointMatrixIndexProvider rows match. */
   static MultiBodySystemBasics build(Exchange x)
   {
      RigidBodyBasics elevator = new RigidBody("elevator", ReferenceFrame.getWorldFrame()); // multiBodySystem/RigidBody.java:79
      RigidBodyBasics[] successor = new RigidBodyBasics[x.nb];
      for (int i = 0; i < x.nb; i++)
      {
         RigidBodyBasics predecessor = x.parent[i] < 0 ? elevator : successor[x.parent[i]];
         RigidBodyTransform offset = transform(x.offsetR[i], x.offsetP[i]);
         Vector3D axis = new Vector3D(x.axis[i][0], x.axis[i][1], x.axis[i][2]);
         JointBasics joint;
         if (x.jointType[i] == 0)
            joint = new RevoluteJoint("joint" + i, predecessor, offset, axis); // RevoluteJoint.java:69
         else if (x.jointType[i] == 1)
            joint = new PrismaticJoint("joint" + i, predecessor, offset, axis); // PrismaticJoint.java:47
         else
            joint = new SixDoFJoint("joint" + i, predecessor, offset); // SixDoFJoint.java:64
         double[] J = x.inertia[i];
         successor[i] = new RigidBody("body" + i, joint, new Matrix3D(J[0], J[1], J[2], J[3], J[4], J[5], J[6], J[7], J[8]), x.mass[i],
                                      transform(x.comR[i], x.comP[i])); // RigidBody.java:163
      }
      return MultiBodySystemBasics.toMultiBodySystemBasics(elevator); // MultiBodySystemBasics.java:76
   }

   static void column(double[] rows, int nRows, int n, int s, DMatrixRMaj out)
   {
      out.reshape(nRows, 1);
      for (int r = 0; r < nRows; r++)
         out.set(r, 0, rows[r * n + s]);
   }

   /** One worker: its own system + calculators (they are not thread-safe, MultiBodySystemFactories.java:310-348). */
   static class Worker
   {
      final Exchange x;
      final MultiBodySystemBasics system;
      final List<? extends JointBasics> joints;
      final InverseDynamicsCalculator inverseDynamics;
      final ForwardDynamicsCalculator forwardDynamics;
      final CompositeRigidBodyMassMatrixCalculator massMatrix;
      final DMatrixRMaj q = new DMatrixRMaj(1, 1), qd = new DMatrixRMaj(1, 1), qdd = new DMatrixRMaj(1, 1), tau = new DMatrixRMaj(1, 1);

      Worker(Exchange x)
      {
         this.x = x;
         system = build(x);
         joints = system.getJointsToConsider();
         inverseDynamics = new InverseDynamicsCalculator(system);          // InverseDynamicsCalculator.java:201
         forwardDynamics = new ForwardDynamicsCalculator(system);          // ForwardDynamicsCalculator.java:128
         massMatrix = new CompositeRigidBodyMassMatrixCalculator(system);  // CompositeRigidBodyMassMatrixCalculator.java:182
         inverseDynamics.setGravitationalAcceleration(x.gravity[0], x.gravity[1], x.gravity[2]); // :397
         forwardDynamics.setGravitationalAcceleration(x.gravity[0], x.gravity[1], x.gravity[2]); // :313
      }

      /** The hot path of SURVEY.md section 3 for state s: insert state -> update frames -> the three calculators. */
      void evaluate(int s, double[] tauOut, double[] qddOut, double[] massOut)
      {
         column(x.q, x.nq, x.n, s, q); column(x.qd, x.nv, x.n, s, qd); column(x.qdd, x.nv, x.n, s, qdd); column(x.tau, x.nv, x.n, s, tau);
         MultiBodySystemTools.insertJointsState(joints, JointStateType.CONFIGURATION, q); // tools/MultiBodySystemTools.java:1578
         MultiBodySystemTools.insertJointsState(joints, JointStateType.VELOCITY, qd);
         system.getRootBody().updateFramesRecursively();
         inverseDynamics.compute(qdd);                                       // :496
         DMatrixRMaj t = inverseDynamics.getJointTauMatrix();                // :567
         forwardDynamics.compute(tau);                                       // :489
         DMatrixRMaj a = forwardDynamics.getJointAccelerationMatrix();      // :556
         massMatrix.reset();                                                 // :286
         DMatrixRMaj M = massMatrix.getMassMatrix();                         // :344
         if (tauOut != null)
         {
            for (int r = 0; r < x.nv; r++)
            {
               tauOut[r * x.n + s] = t.get(r, 0);
               qddOut[r * x.n + s] = a.get(r, 0);
               for (int c = 0; c < x.nv; c++)
                  massOut[(r * x.nv + c) * x.n + s] = M.get(r, c);
            }
         }
      }
   }

   static void dump(String in, String out) throws IOException
   {
      Exchange x = read(in);
      Worker w = new Worker(x);
      double[] tau = new double[x.nv * x.n], qdd = new double[x.nv * x.n], M = new double[x.nv * x.nv * x.n];
      for (int s = 0; s < x.n; s++)
         w.evaluate(s, tau, qdd, M);
      ByteBuffer b = ByteBuffer.allocate(8 * (tau.length + qdd.length + M.length)).order(ByteOrder.LITTLE_ENDIAN);
      for (double v : tau) b.putDouble(v);
      for (double v : qdd) b.putDouble(v);
      for (double v : M) b.putDouble(v);
      b.flip();
      try (FileChannel ch = FileChannel.open(Paths.get(out), StandardOpenOption.CREATE, StandardOpenOption.WRITE, StandardOpenOption.TRUNCATE_EXISTING))
      {
         while (b.hasRemaining())
            ch.write(b);
      }
      System.out.println("{\"states\": " + x.n + ", \"n_dofs\": " + x.nv + ", \"written\": \"" + out + "\"}");
   }

   static void bench(String in, int threads, double seconds) throws Exception
   {
      Exchange x = read(in);
      List<Thread> pool = new ArrayList<>();
      AtomicLong states = new AtomicLong();
      long warmup = 5000; // JIT warm-up like the reference's own harness (InverseDynamicsCalculatorTest.java:21-22)
      long[] t0 = new long[1];
      Object gate = new Object();
      int[] ready = {0};
      for (int k = 0; k < threads; k++)
      {
         final int id = k;
         Thread th = new Thread(() ->
         {
            Worker w = new Worker(x);
            int s = id % x.n;
            for (long i = 0; i < warmup; i++, s = (s + threads) % x.n)
               w.evaluate(s, null, null, null);
            synchronized (gate)
            {
               if (++ready[0] == threads)
               {
                  t0[0] = System.nanoTime();
                  gate.notifyAll();
               }
               else
                  while (ready[0] < threads)
                     try { gate.wait(); } catch (InterruptedException e) { return; }
            }
            long mine = 0;
            while ((System.nanoTime() - t0[0]) * 1e-9 < seconds)
            {
               for (int i = 0; i < 64; i++, s = (s + threads) % x.n)
                  w.evaluate(s, null, null, null);
               mine += 64;
            }
            states.addAndGet(mine);
         });
         th.start();
         pool.add(th);
      }
      for (Thread th : pool)
         th.join();
      double elapsed = (System.nanoTime() - t0[0]) * 1e-9;
      System.out.println("{\"metric\": \"RNEA+ABA+CRBA states/sec\", \"impl\": \"mecano-jvm\", \"value\": " + states.get() / elapsed + ", \"unit\": \"states/s\", \"cores\": "
            + threads + ", \"seconds\": " + elapsed + ", \"n_dofs\": " + x.nv + "}");
   }

   public static void main(String[] args) throws Exception
   {
      if (args.length >= 3 && args[0].equals("dump"))
         dump(args[1], args[2]);
      else if (args.length >= 2 && args[0].equals("bench"))
         bench(args[1], args.length > 2 ? Integer.parseInt(args[2]) : Runtime.getRuntime().availableProcessors(),
               args.length > 3 ? Double.parseDouble(args[3]) : 10.0);
      else
         System.err.println("usage: MecanoHarness dump <exchange.bin> <results.bin> | bench <exchange.bin> [threads] [seconds]");
   }
}
