/*
 * Pins the CPU oracle (oracle/*.c) and the CUDA path to the REAL sdrtrunk classes.  The reference ships no golden vectors
 * for its DSP path and the build image of this repository has no JVM, so oracle-vs-Java parity is "unpinned" there; this
 * harness is what closes that gap on any machine with a JDK and the sdrtrunk classes on the class path:
 *
 *     tools/mint_jvm_goldens.sh /path/to/sdrtrunk            (exports inputs, runs this class, imports its outputs)
 *
 * It reads little-endian float32 files from <dir>/in, drives the unmodified reference classes -- wired the way the
 * decoders wire them (P25P1DecoderC4FM.java:62-116, P25P1DecoderLSM.java:198-236, P25P2DecoderHDQPSK.java:249-305,
 * NBFMDecoder.java:129-181) -- and writes what they produce to <dir>/out.  tests/test_jvm_goldens.py then requires the
 * oracle and the CUDA path to reproduce those files (bit for bit where SURVEY.md 8(c) says so).
 */
package io.github.dsheirer.gpu;

import io.github.dsheirer.dsp.filter.FilterFactory;
import io.github.dsheirer.dsp.filter.channelizer.ComplexPolyphaseChannelizerM2;
import io.github.dsheirer.dsp.filter.channelizer.output.OneChannelOutputProcessor;
import io.github.dsheirer.dsp.filter.channelizer.output.TwoChannelOutputProcessor;
import io.github.dsheirer.dsp.filter.decimate.DecimationFilterFactory;
import io.github.dsheirer.dsp.filter.decimate.IComplexDecimationFilter;
import io.github.dsheirer.dsp.filter.fir.FIRFilterSpecification;
import io.github.dsheirer.dsp.filter.fir.complex.ComplexFIRFilter2;
import io.github.dsheirer.dsp.fm.FMDemodulator;
import io.github.dsheirer.dsp.fm.SquelchingFMDemodulator;
import io.github.dsheirer.dsp.gain.ComplexFeedForwardGainControl;
import io.github.dsheirer.dsp.psk.DQPSKDecisionDirectedDemodulator;
import io.github.dsheirer.dsp.psk.DQPSKGardnerDemodulator;
import io.github.dsheirer.dsp.psk.InterpolatingSampleBuffer;
import io.github.dsheirer.dsp.psk.PSKDemodulator;
import io.github.dsheirer.dsp.psk.pll.CostasLoop;
import io.github.dsheirer.dsp.psk.pll.PLLBandwidth;
import io.github.dsheirer.dsp.symbol.Dibit;
import io.github.dsheirer.sample.buffer.ReusableChannelResultsBuffer;
import io.github.dsheirer.sample.buffer.ReusableComplexBuffer;
import io.github.dsheirer.sample.buffer.ReusableComplexBufferAssembler;
import io.github.dsheirer.sample.buffer.ReusableComplexBufferQueue;
import io.github.dsheirer.sample.buffer.ReusableFloatBuffer;

import java.io.ByteArrayOutputStream;
import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.file.Files;
import java.nio.file.Path;
import java.util.ArrayList;
import java.util.Arrays;
import java.util.Collections;
import java.util.List;

public class OracleHarness
{
    private static final int BUFFER_FLOATS = 2048;      //PolyphaseChannelSource.java:42: 1024 complex samples per buffer
    private final Path mIn;
    private final Path mOut;
    private final ReusableComplexBufferQueue mQueue = new ReusableComplexBufferQueue("OracleHarness");

    public OracleHarness(Path directory) throws IOException
    {
        mIn = directory.resolve("in");
        mOut = directory.resolve("out");
        Files.createDirectories(mOut);
    }

    public static void main(String[] args) throws Exception
    {
        OracleHarness harness = new OracleHarness(Path.of(args.length > 0 ? args[0] : "tests/golden_jvm"));
        harness.remez();
        harness.channelizer();
        harness.filters();
        harness.fm();
        harness.p25("c4fm", 4800.0, PLLBandwidth.BW_300, 0.3f, false, harness.c4fmTaps());
        harness.p25("dmr", 4800.0, PLLBandwidth.BW_300, 0.4f, false, harness.c4fmTaps());
        harness.p25("lsm", 4800.0, PLLBandwidth.BW_200, 0.3f, true, null);
        harness.p25("hdqpsk", 6000.0, PLLBandwidth.BW_300, 0.1f, true, harness.hdqpskTaps());
        System.out.println("OracleHarness: wrote " + harness.mOut.toAbsolutePath());
        System.exit(0);
    }

    // ------------------------------------------------------------------------------------------------ file helpers
    private float[] read(String name) throws IOException
    {
        ByteBuffer bytes = ByteBuffer.wrap(Files.readAllBytes(mIn.resolve(name))).order(ByteOrder.LITTLE_ENDIAN);
        float[] values = new float[bytes.remaining() / 4];
        bytes.asFloatBuffer().get(values);
        return values;
    }

    private void write(String name, float[] values) throws IOException
    {
        ByteBuffer bytes = ByteBuffer.allocate(4 * values.length).order(ByteOrder.LITTLE_ENDIAN);
        bytes.asFloatBuffer().put(values);
        Files.write(mOut.resolve(name), bytes.array());
    }

    private static float[] concat(List<float[]> parts)
    {
        int total = 0;

        for(float[] part : parts)
        {
            total += part.length;
        }

        float[] out = new float[total];
        int pointer = 0;

        for(float[] part : parts)
        {
            System.arraycopy(part, 0, out, pointer, part.length);
            pointer += part.length;
        }

        return out;
    }

    /** a reusable buffer holding a copy of samples[from, from + length), user count 1 as the sources hand them out */
    private ReusableComplexBuffer buffer(float[] samples, int from, int length)
    {
        ReusableComplexBuffer buffer = mQueue.getBuffer(Arrays.copyOfRange(samples, from, from + length), 0L);
        buffer.incrementUserCount();
        return buffer;
    }

    // ------------------------------------------------------------------------------------------------ Remez designer
    private float[] c4fmTaps() throws Exception
    {
        //P25P1DecoderC4FM.getBasebandFilter (:136-148) at the polyphase channel rate
        return FilterFactory.getTaps(FIRFilterSpecification.lowPassBuilder().sampleRate(50000).passBandCutoff(5100)
            .passBandAmplitude(1.0).passBandRipple(0.01).stopBandAmplitude(0.0).stopBandStart(6500).stopBandRipple(0.01).build());
    }

    private float[] hdqpskTaps() throws Exception
    {
        //P25P2DecoderHDQPSK.getBasebandFilter (:155-166)
        return FilterFactory.getTaps(FIRFilterSpecification.lowPassBuilder().sampleRate(50000.0).passBandCutoff(6500)
            .passBandAmplitude(1.0).passBandRipple(0.005).stopBandAmplitude(0.0).stopBandStart(7200).stopBandRipple(0.01).build());
    }

    private float[] nbfmTaps() throws Exception
    {
        //NBFMDecoder (:306-325) for a 12.5 kHz channel at the decimated rate of 25 kHz
        return FilterFactory.getTaps(FIRFilterSpecification.lowPassBuilder().sampleRate(25000.0 * 2).gridDensity(16).oddLength(true)
            .passBandCutoff(10000).passBandAmplitude(1.0).passBandRipple(0.01).stopBandStart(12500).stopBandAmplitude(0.0)
            .stopBandRipple(0.005).build());
    }

    private void remez() throws Exception
    {
        write("remez_c4fm.f32", c4fmTaps());
        write("remez_hdqpsk.f32", hdqpskTaps());
        write("remez_nbfm.f32", nbfmTaps());
    }

    // ------------------------------------------------------------------------------------------------ channelizer
    /** the real channelizer with dispatch() intercepted: channel results come back on the IFFT processor's thread */
    private static class CapturingChannelizer extends ComplexPolyphaseChannelizerM2
    {
        final List<float[]> mResults = Collections.synchronizedList(new ArrayList<>());

        CapturingChannelizer(float[] taps, int sampleRate, int channelCount)
        {
            super(taps, sampleRate, channelCount);
        }

        @Override
        protected void dispatch(ReusableChannelResultsBuffer channelResultsBuffer)
        {
            for(float[] results : channelResultsBuffer.getChannelResults())
            {
                mResults.add(Arrays.copyOf(results, results.length));
            }

            channelResultsBuffer.decrementUserCount();
        }
    }

    private void channelizer() throws Exception
    {
        int m = 96;
        float[] x = read("channelizer_m96_x.f32");
        float[] taps = FilterFactory.getSincM2Channelizer(25000.0, m, 9, false);
        write("channelizer_m96_taps.f32", taps);

        CapturingChannelizer channelizer = new CapturingChannelizer(taps, 25000 * m, m);
        channelizer.start();
        channelizer.receive(buffer(x, 0, x.length));
        int blocks = x.length / m;     //one block per M / 2 complex samples = M floats

        for(int wait = 0; wait < 200 && channelizer.mResults.size() < blocks; wait++)
        {
            Thread.sleep(50);
        }

        channelizer.stop();
        List<float[]> results = new ArrayList<>(channelizer.mResults);
        write("channelizer_m96_results.f32", concat(results));

        //OneChannelOutputProcessor on bin 4 with a +700 Hz correction; TwoChannelOutputProcessor on bins 88 + 89, -300 Hz:
        //the amplitude convention of the two-bin path (SURVEY.md a7) is whatever these classes produce
        float[] synthesis = FilterFactory.getSincM2Synthesizer(50000.0, 25000.0, 2, 9);
        write("synth.f32", synthesis);
        write("bin4_offset700.f32", outputProcessor(new OneChannelOutputProcessor(50000.0, List.of(4), m), 700, results));
        write("bins88_89_offset_m300.f32",
            outputProcessor(new TwoChannelOutputProcessor(50000.0, List.of(88, 89), synthesis, m), -300, results));
    }

    private float[] outputProcessor(io.github.dsheirer.dsp.filter.channelizer.output.ChannelOutputProcessor processor, long offset,
                                    List<float[]> results)
    {
        processor.setFrequencyOffset(offset);
        List<float[]> out = new ArrayList<>();
        ReusableComplexBufferAssembler assembler = new ReusableComplexBufferAssembler(BUFFER_FLOATS, 50000.0);
        assembler.setListener(buffer -> {
            out.add(buffer.getSamplesCopy());
            buffer.decrementUserCount();
        });
        io.github.dsheirer.sample.buffer.ReusableChannelResultsBufferQueue queue =
            new io.github.dsheirer.sample.buffer.ReusableChannelResultsBufferQueue("OracleHarness");
        ReusableChannelResultsBuffer resultsBuffer = queue.getBuffer();

        for(float[] block : results)
        {
            resultsBuffer.addChannelResults(Arrays.copyOf(block, block.length));
        }

        resultsBuffer.incrementUserCount();
        processor.process(List.of(resultsBuffer), assembler);
        assembler.flush();
        return concat(out);
    }

    // ------------------------------------------------------------------------------------------------ filters, AGC
    private void filters() throws Exception
    {
        float[] x = read("filters_x.f32");        //two 1024-sample buffers
        IComplexDecimationFilter decimator = DecimationFilterFactory.getComplexDecimationFilter(8);
        ComplexFIRFilter2 fir = new ComplexFIRFilter2(c4fmTaps());
        ComplexFeedForwardGainControl agc = new ComplexFeedForwardGainControl(32);   //P25P1Decoder: window of 32
        List<float[]> decimated = new ArrayList<>(), filtered = new ArrayList<>(), gained = new ArrayList<>();

        for(int from = 0; from + BUFFER_FLOATS <= x.length; from += BUFFER_FLOATS)
        {
            ReusableComplexBuffer a = decimator.decimate(buffer(x, from, BUFFER_FLOATS));
            decimated.add(a.getSamplesCopy());
            a.decrementUserCount();
            ReusableComplexBuffer b = fir.filter(buffer(x, from, BUFFER_FLOATS));
            filtered.add(b.getSamplesCopy());
            b.decrementUserCount();
            ReusableComplexBuffer c = agc.filter(buffer(x, from, BUFFER_FLOATS));
            gained.add(c.getSamplesCopy());
            c.decrementUserCount();
        }

        write("decimate8.f32", concat(decimated));
        write("fir72.f32", concat(filtered));
        write("agc.f32", concat(gained));
    }

    // ------------------------------------------------------------------------------------------------ FM
    private void fm() throws Exception
    {
        float[] x = read("fm_x.f32");
        FMDemodulator demodulator = new FMDemodulator(1.0f);
        SquelchingFMDemodulator squelching = new SquelchingFMDemodulator(0.01, -40.0, 4);
        ReusableFloatBuffer a = demodulator.demodulate(buffer(x, 0, x.length));
        write("fm.f32", a.getSamplesCopy());
        a.decrementUserCount();
        ReusableFloatBuffer b = squelching.demodulate(buffer(x, 0, x.length));
        write("squelch_fm.f32", b.getSamplesCopy());
        b.decrementUserCount();

        //NBFMDecoder.receive (:129-181): decimate by 2 -> I/Q filter -> squelching demodulator, per 1024-sample buffer
        float[] wide = read("nbfm_x.f32");
        IComplexDecimationFilter decimator = DecimationFilterFactory.getComplexDecimationFilter(2);
        ComplexFIRFilter2 iqFilter = new ComplexFIRFilter2(nbfmTaps());
        SquelchingFMDemodulator nbfm = new SquelchingFMDemodulator(0.0004, -78.0, 4);    //NBFMDecoder.java:55-62
        List<float[]> audio = new ArrayList<>();

        for(int from = 0; from + BUFFER_FLOATS <= wide.length; from += BUFFER_FLOATS)
        {
            ReusableFloatBuffer demodulated = nbfm.demodulate(iqFilter.filter(decimator.decimate(buffer(wide, from, BUFFER_FLOATS))));
            audio.add(demodulated.getSamplesCopy());
            demodulated.decrementUserCount();
        }

        write("nbfm_audio.f32", concat(audio));
    }

    // ------------------------------------------------------------------------------------------------ P25 / DMR chains
    /**
     * baseband filter -> ComplexFeedForwardGainControl -> DQPSK demodulator with CostasLoop + InterpolatingSampleBuffer, as
     * P25P1DecoderC4FM.setSampleRate / receive (:68-116), P25P1DecoderLSM (:198-236), P25P2DecoderHDQPSK (:249-305) and
     * DMRDecoder (:58-131) build and run them, with a Listener<Dibit> on the demodulator and no message framer behind it
     * (so no PLL inversion feedback: SURVEY.md 8d config 3).
     */
    private void p25(String name, double symbolRate, PLLBandwidth bandwidth, float sampleCounterGain, boolean gardner,
                     float[] basebandTaps) throws Exception
    {
        double sampleRate = 50000.0;
        float[] x = read(name + "_x.f32");
        ComplexFIRFilter2 basebandFilter = basebandTaps != null ? new ComplexFIRFilter2(basebandTaps) : null;
        ComplexFeedForwardGainControl agc = new ComplexFeedForwardGainControl(32);
        CostasLoop costasLoop = new CostasLoop(sampleRate, symbolRate);
        costasLoop.setPLLBandwidth(bandwidth);
        InterpolatingSampleBuffer sampleBuffer = new InterpolatingSampleBuffer((float)(sampleRate / symbolRate), sampleCounterGain);
        PSKDemodulator<Dibit> demodulator = gardner ? new DQPSKGardnerDemodulator(costasLoop, sampleBuffer)
            : new DQPSKDecisionDirectedDemodulator(costasLoop, sampleBuffer);
        ByteArrayOutputStream dibits = new ByteArrayOutputStream();
        demodulator.setSymbolListener(dibit -> dibits.write(dibit.getValue()));
        List<float[]> gained = new ArrayList<>();

        for(int from = 0; from + BUFFER_FLOATS <= x.length; from += BUFFER_FLOATS)
        {
            ReusableComplexBuffer filtered = buffer(x, from, BUFFER_FLOATS);

            if(basebandFilter != null)
            {
                filtered = basebandFilter.filter(filtered);
            }

            ReusableComplexBuffer gainApplied = agc.filter(filtered);
            gained.add(gainApplied.getSamplesCopy());
            demodulator.receive(gainApplied);
        }

        write(name + "_agc.f32", concat(gained));
        Files.write(mOut.resolve(name + "_dibits.u8"), dibits.toByteArray());
        write(name + "_loop.f32", new float[]{(float)costasLoop.getLoopFrequency(), sampleBuffer.getSamplingPoint()});
    }
}
